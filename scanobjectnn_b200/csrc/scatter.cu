// scatter.cu -- the three scatter-add gradients of the reference's custom ops, without floating-point atomics:
//   GroupPointGrad        tf_grouping_g.cu:61-78   (atomicAdd in the reference's CUDA kernel)
//   GatherPointGrad       tf_sampling_g.cu:183-192 (atomicAdd)
//   ThreeInterpolateGrad  tf_interpolate.cpp:131-153 (sequential CPU loop)
// All three are "dst[idx[e]] += w[e] * src[e / div]" over the entries e of one cloud.  Two kernels:
//  1. group_csr_kernel (one CTA per cloud): a STABLE counting sort of the cloud's entries by destination: counts (shared-memory
//     integer atomics: order-free), exclusive scan, then one warp walks the entries in order, 32 at a time -- __match_any_sync
//     groups the lanes that reference the same destination, the rank inside the group is a popcount of the lower lanes, the
//     group leader advances that destination's cursor -- so list[] holds, per destination, its entries in ascending order;
//  2. scatter_rows_kernel (thread = one (destination, channel) pair): adds the entries of the destination in list order.
// The sum order is therefore fixed AND equal to the order of the reference's sequential CPU loops (ascending entry): results are
// bit-reproducible run to run and bit-identical to oracle/psa_oracle.c (orc_group_point_grad, orc_gather_point_grad,
// orc_three_interpolate_grad: products rounded before the add, as the x86 build without FMA evaluates them).
#include "common.cuh"

namespace psa {

__global__ void __launch_bounds__(256) group_csr_kernel(int n, int mk, const int* __restrict__ idx, int* __restrict__ offsets, int* __restrict__ list) {
    extern __shared__ int sm_i[];                 // n counters / cursors
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int* ic = idx + (size_t)cloud * mk;
    int* off = offsets + (size_t)cloud * (n + 1);
    int* lst = list + (size_t)cloud * mk;
    for (int j = tid; j < n; j += 256) sm_i[j] = 0;
    __syncthreads();
    for (int e = tid; e < mk; e += 256) {
        const int v = __ldg(ic + e);
        if ((unsigned)v < (unsigned)n) atomicAdd(&sm_i[v], 1);                 // out-of-range entries are dropped, not scattered
    }
    __syncthreads();
    // exclusive scan by one warp (n <= a few thousand): lane owns a contiguous run of counters
    if (tid < 32) {
        const int per = (n + 31) / 32;
        const int j0 = min(n, lane * per), j1 = min(n, j0 + per);
        int local = 0;
        for (int j = j0; j < j1; ++j) local += sm_i[j];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - local;
        for (int j = j0; j < j1; ++j) { const int c = sm_i[j]; sm_i[j] = run; off[j] = run; run += c; }
        if (lane == 31) off[n] = incl;
        __syncwarp();
        // stable placement: entries in order, 32 per step
        for (int base = 0; base < mk; base += 32) {
            const int e = base + lane;
            int v = e < mk ? __ldg(ic + e) : -1;
            const bool act = (unsigned)v < (unsigned)n;
            if (!act) v = -1 - lane;                                           // inactive lanes: distinct dummies
            const unsigned grp = __match_any_sync(0xffffffffu, v);
            const int rank = __popc(grp & lanemask_lt());
            int cur = 0;
            if (act) cur = sm_i[v];
            __syncwarp();
            if (act) {
                lst[cur + rank] = e;
                if (rank == 0) sm_i[v] = cur + __popc(grp);                     // the group's lowest lane advances the cursor
            }
            __syncwarp();
        }
    }
}

int launch_group_csr(int b, int n, int mk, const int* idx, int* offsets, int* list, cudaStream_t st) {
    PSA_SUPPORTED((size_t)n * sizeof(int) <= 200 * 1024, "scatter-add: %d destinations per cloud exceed the shared-memory counters", n);
    const size_t smem = (size_t)n * sizeof(int);
    PSA_CUDA(cudaFuncSetAttribute(group_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_csr_kernel<<<b, 256, smem, st>>>(n, mk, idx, offsets, list);
    return check_launch("group_csr_kernel");
}

namespace {

// thread = (destination point, channel), channel fastest: the threads of one destination read each source row coalesced and
// the list entry as a broadcast.  Four entries in flight; the adds stay in list order.
template <bool WEIGHTED>
__global__ void __launch_bounds__(256)
scatter_rows_kernel(int n, int mk, int div, int c, long long total, const float* __restrict__ src, const float* __restrict__ weight,
                    const int* __restrict__ offsets, const int* __restrict__ list, float* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const long long pt = i / c;
    const int l = (int)(i - pt * c);
    const long long cloud = pt / n;
    const int j = (int)(pt - cloud * n);
    const int* off = offsets + (size_t)cloud * (n + 1);
    const int* lst = list + (size_t)cloud * mk;
    const float* sc = src + (size_t)cloud * (mk / div) * c + l;
    const float* wc = WEIGHTED ? weight + (size_t)cloud * mk : nullptr;
    const int t0 = __ldg(off + j), t1 = __ldg(off + j + 1);
    float acc = 0.f;
    int t = t0;
    for (; t + 3 < t1; t += 4) {
        int e[4];
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = __ldg(lst + t + u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = __ldg(sc + (size_t)(e[u] / div) * c);
            if (WEIGHTED) v[u] = __fmul_rn(v[u], __ldg(wc + e[u]));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = __fadd_rn(acc, v[u]);
    }
    for (; t < t1; ++t) {
        const int e = __ldg(lst + t);
        float v = __ldg(sc + (size_t)(e / div) * c);
        if (WEIGHTED) v = __fmul_rn(v, __ldg(wc + e));
        acc = __fadd_rn(acc, v);
    }
    dst[i] = acc;
}

size_t scatter_ws_bytes(int b, int n_dst, long long entries) {
    return ((size_t)b * ((size_t)n_dst + 1) + (size_t)b * (size_t)entries) * sizeof(int) + 256;
}

// dst (b, n_dst, c) = ordered scatter-add of src (b, entries / div, c) rows through idx (b, entries) [x weight (b, entries)]
int scatter_add_ordered(const char* what, int b, int n_dst, long long entries, int div, int c, const float* src, const int* idx,
                        const float* weight, float* dst, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const long long total = (long long)b * n_dst * c;
    if (total == 0) return PSA_OK;
    PSA_REQUIRE(dst != nullptr, "%s: null output", what);
    if (entries == 0) {
        PSA_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)total, st));
        return PSA_OK;
    }
    PSA_REQUIRE(src && idx, "%s: null buffer", what);
    PSA_SUPPORTED(entries <= 0x7fffffffLL, "%s: %lld entries per cloud exceed int32", what, entries);
    const size_t need = scatter_ws_bytes(b, n_dst, entries);
    PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, "%s: workspace of %zu bytes needed (psa_scatter_workspace_bytes), got %zu", what,
                need, workspace_bytes);
    int* offsets = reinterpret_cast<int*>(workspace);
    int* list = offsets + (size_t)b * ((size_t)n_dst + 1);
    int rc = launch_group_csr(b, n_dst, (int)entries, idx, offsets, list, st);
    if (rc != PSA_OK) return rc;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (weight)
        scatter_rows_kernel<true><<<blocks, 256, 0, st>>>(n_dst, (int)entries, div, c, total, src, weight, offsets, list, dst);
    else
        scatter_rows_kernel<false><<<blocks, 256, 0, st>>>(n_dst, (int)entries, div, c, total, src, nullptr, offsets, list, dst);
    return check_launch("scatter_rows_kernel");
}

}  // namespace
}  // namespace psa

using namespace psa;

extern "C" size_t psa_scatter_workspace_bytes(int b, int n_dst, long long entries) {
    if (b <= 0 || n_dst < 0 || entries <= 0) return 256;
    return scatter_ws_bytes(b, n_dst, entries);
}

extern "C" int psa_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                                    float* grad_points, void* workspace, size_t workspace_bytes, psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 0 && c >= 0 && m >= 0 && nsample >= 0, "GroupPointGrad: negative dimension");
    return scatter_add_ordered("GroupPointGrad", b, n, (long long)m * nsample, 1, c, grad_out, idx, nullptr, grad_points, workspace,
                               workspace_bytes, as_stream(stream));
}

extern "C" int psa_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g, void* workspace,
                                     size_t workspace_bytes, psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "GatherPointGrad: negative dimension");
    return scatter_add_ordered("GatherPointGrad", b, n, m, 1, 3, out_g, idx, nullptr, inp_g, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int psa_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight,
                                          float* grad_points, void* workspace, size_t workspace_bytes, psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 0 && m >= 0 && c >= 0, "ThreeInterpolateGrad: negative dimension");
    PSA_REQUIRE(weight != nullptr || (long long)b * n * c == 0, "ThreeInterpolateGrad: null weight");
    return scatter_add_ordered("ThreeInterpolateGrad", b, m, 3LL * n, 3, c, grad_out, idx, weight, grad_points, workspace, workspace_bytes,
                               as_stream(stream));
}

// sa_train.cu -- training-mode front of a set-abstraction level ("variant F1", SURVEY 7 hard part 4).
//
// In training the reference's relu(BN(conv(x)+b)) needs batch statistics over all B*m*K rows before the ReLU
// (pointnet2/utils/tf_util.py:512-531, is_training=True), so the first 1x1 conv's PRE-BN output has to exist in HBM
// once.  The reference gets there through query_ball_point -> group_point -> tile/sub -> concat -> cuDNN conv + bias_add:
// four materialised (B,m,K,.) tensors.  This kernel does the whole front in ONE launch:
//   ball query (index-exact, grouping.cu's warp scan) -> neighbour coordinates straight from the shared-memory copy of
//   the cloud -> centre -> conv1 (+ optional per-point feature products U = points . W1[3:,:]) + bias ->
//   coalesced streaming store of (B,m,K,C1) + idx/pts_cnt + per-channel sum / sum-of-squares for the BN statistics.
// HBM traffic = algorithmic traffic: B*(12n + 12m) in, 4*B*m*K*(C1+1) + 4*B*m out  (137.3 MB at B=32,N=2048,m=512,K=32,
// C1=64) -- a pure write-bound kernel, the one BASELINE.json's ">= 70 % of the HBM roofline" target is defined on.
#include "ball_query.cuh"
#include "common.cuh"

namespace psa {

float ball_query_threshold(float radius, bool* none);   // grouping.cu

constexpr int kF1Warps = kBqWarps;

struct F1Args {
    int n, m, nsample, C1, q_per_cta;
    float radius, thr;
    int none, want_grid;
    const float* xyz;      // (b,n,3)
    const float* new_xyz;  // (b,m,3)
    const float* uf;       // (b*n, C1) or null
    const float* w1;       // (3+c, C1): rows 0..2 used here
    const float* bias;     // (C1) or null
    float* pre;            // (b,m,K,C1)
    int* idx;              // (b,m,K)
    int* pts_cnt;          // (b,m) or null
    float* partial;        // (gridDim.x*gridDim.y, 2, C1) or null
};

// conv stage mapping: 8 lanes per row, 4 rows per warp step; lane-in-row s owns channels [32*i + 4*s, +4), i < NV = C1/32,
// so every store instruction writes 128 contiguous bytes per row and a query's K rows take K/4 steps.
template <int NV, bool STATS, int PPT>
__global__ void __launch_bounds__(kF1Warps * 32, PPT <= 8 ? 3 : 2)
sa_conv1_prebn_kernel(const __grid_constant__ F1Args a) {
    extern __shared__ __align__(16) float smem_f[];
    const int n = a.n;
    const int cloud = blockIdx.y;
    const float* gx = a.xyz + (size_t)cloud * n * 3;       // neighbour coordinates come from global memory / L1 (24 KB per cloud)
    const BqSmem s = bq_carve(smem_f, n, a.want_grid != 0, gx);
    int* srow = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(smem_f) + bq_smem_bytes(n, a.want_grid != 0));   // kF1Warps * nsample
    float* sstat = reinterpret_cast<float*>(srow + kF1Warps * ((a.nsample + 3) & ~3));   // kF1Warps * 2 * C1 (STATS), 16-B aligned
    const BqGrid g = bq_stage_and_build<PPT>(s, n, a.radius, a.want_grid != 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, rsub = lane >> 3;
    // packed f32x2 registers: pair p of vector i covers channels 32*i + 4*sub + 2*p, +1
    constexpr int NPK = 2 * NV;
    float2 wx[NPK], wy[NPK], wz[NPK], bs[NPK], ssum[NPK], ssq[NPK];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 32 * i + 4 * sub;
        const float4 x4 = __ldg(reinterpret_cast<const float4*>(a.w1 + c));
        const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.w1 + a.C1 + c));
        const float4 z4 = __ldg(reinterpret_cast<const float4*>(a.w1 + 2 * a.C1 + c));
        const float4 b4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        wx[2 * i] = make_float2(x4.x, x4.y); wx[2 * i + 1] = make_float2(x4.z, x4.w);
        wy[2 * i] = make_float2(y4.x, y4.y); wy[2 * i + 1] = make_float2(y4.z, y4.w);
        wz[2 * i] = make_float2(z4.x, z4.y); wz[2 * i + 1] = make_float2(z4.z, z4.w);
        bs[2 * i] = make_float2(b4.x, b4.y); bs[2 * i + 1] = make_float2(b4.z, b4.w);
        ssum[2 * i] = ssum[2 * i + 1] = make_float2(0.f, 0.f);
        ssq[2 * i] = ssq[2 * i + 1] = make_float2(0.f, 0.f);
    }
    const int q0 = blockIdx.x * a.q_per_cta;
    const int q1 = min(a.m, q0 + a.q_per_cta);
    const float* p2 = a.new_xyz + (size_t)cloud * a.m * 3;
    int* row = srow + warp * a.nsample;
    const float* ucloud = a.uf ? a.uf + (size_t)cloud * n * a.C1 + 4 * sub : nullptr;
    for (int q = q0 + warp; q < q1; q += kF1Warps) {
        const float qx = __ldg(p2 + q * 3 + 0), qy = __ldg(p2 + q * 3 + 1), qz = __ldg(p2 + q * 3 + 2);
        const int cnt = bq_query_warp(n, a.nsample, a.thr, a.none != 0, s, g, qx, qy, qz, row, lane, warp);
        __syncwarp();
        const size_t gq = (size_t)cloud * a.m + q;
        for (int l = lane; l < a.nsample; l += 32) a.idx[gq * a.nsample + l] = row[l];
        if (a.pts_cnt != nullptr && lane == 0) a.pts_cnt[gq] = cnt;
        float* outl = a.pre + gq * a.nsample * a.C1 + 4 * sub;          // this lane's column offset inside the query's block
        for (int r0 = 0; r0 < a.nsample; r0 += 4) {
            const int r = r0 + rsub;
            if (r < a.nsample) {
                const int j = row[r];
                // grouped_xyz - new_xyz (pointnet_util.py:46), broadcast into both halves of a packed register
                const float dxs = __ldg(gx + 3 * j) - qx, dys = __ldg(gx + 3 * j + 1) - qy, dzs = __ldg(gx + 3 * j + 2) - qz;
                const float2 dx = make_float2(dxs, dxs), dy = make_float2(dys, dys), dz = make_float2(dzs, dzs);
                const float* urow = ucloud ? ucloud + (unsigned)j * (unsigned)a.C1 : nullptr;
                float* orow = outl + (unsigned)r * (unsigned)a.C1;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    float2 s0 = bs[2 * i], s1 = bs[2 * i + 1];
                    if (urow) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(urow + 32 * i));
                        s0 = __fadd2_rn(s0, make_float2(u.x, u.y)); s1 = __fadd2_rn(s1, make_float2(u.z, u.w));
                    }
                    // FFMA2: two channels per instruction
                    const float2 v0 = __ffma2_rn(dz, wz[2 * i], __ffma2_rn(dy, wy[2 * i], __ffma2_rn(dx, wx[2 * i], s0)));
                    const float2 v1 = __ffma2_rn(dz, wz[2 * i + 1], __ffma2_rn(dy, wy[2 * i + 1], __ffma2_rn(dx, wx[2 * i + 1], s1)));
                    __stcs(reinterpret_cast<float4*>(orow + 32 * i), make_float4(v0.x, v0.y, v1.x, v1.y));   // streaming store
                    if (STATS) {
                        ssum[2 * i] = __fadd2_rn(ssum[2 * i], v0); ssum[2 * i + 1] = __fadd2_rn(ssum[2 * i + 1], v1);
                        ssq[2 * i] = __ffma2_rn(v0, v0, ssq[2 * i]); ssq[2 * i + 1] = __ffma2_rn(v1, v1, ssq[2 * i + 1]);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (STATS) {
        // the four row-groups of a warp hold partials of the same channels: fold them, then one row-group publishes
#pragma unroll
        for (int p = 0; p < NPK; ++p) {
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
                ssum[p].x += __shfl_xor_sync(0xffffffffu, ssum[p].x, o); ssum[p].y += __shfl_xor_sync(0xffffffffu, ssum[p].y, o);
                ssq[p].x += __shfl_xor_sync(0xffffffffu, ssq[p].x, o); ssq[p].y += __shfl_xor_sync(0xffffffffu, ssq[p].y, o);
            }
            if (rsub == 0) {
                float* w = sstat + (size_t)warp * 2 * a.C1;
                const int c = 32 * (p >> 1) + 4 * sub + 2 * (p & 1);
                *reinterpret_cast<float2*>(w + c) = ssum[p];
                *reinterpret_cast<float2*>(w + a.C1 + c) = ssq[p];
            }
        }
        __syncthreads();
        float* dst = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * a.C1;
        for (int e = threadIdx.x; e < 2 * a.C1; e += blockDim.x) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kF1Warps; ++w) t += sstat[(size_t)w * 2 * a.C1 + e];   // fixed order: deterministic
            dst[e] = t;
        }
    }
}

// stats[0..C1) = sum, stats[C1..2C1) = sum of squares, over all rows; CTA partials added in index order in fp64
__global__ void f1_stats_reduce_kernel(int nparts, int twoC, const float* __restrict__ partial, float* __restrict__ stats) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= twoC) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partial[(size_t)p * twoC + e];
    stats[e] = (float)s;
}

int launch_dense_raw(long long rows, int K, int N, const float* x, const float* W, float* out, cudaStream_t st);   // below

}  // namespace psa

#include "mlp_internal.cuh"

namespace psa {
int launch_dense_raw(long long rows, int K, int N, const float* x, const float* W, float* out, cudaStream_t st) {
    DenseArgs d;
    d.rows = rows; d.K = K; d.N = N; d.pool_k = 1; d.relu = 0;
    d.x = x; d.W = W; d.scale = nullptr; d.shift = nullptr; d.out = out;
    return launch_dense(d, st);
}
static void f1_grid(int b, int m, int* q_per_cta, dim3* grid) {
    int chunks = (2 * kNumSMs + b - 1) / b;
    int q = (m + chunks - 1) / chunks;
    q = ((q + kF1Warps - 1) / kF1Warps) * kF1Warps;
    if (q < kF1Warps) q = kF1Warps;
    *q_per_cta = q;
    *grid = dim3((m + q - 1) / q, b);
}
}  // namespace psa

using namespace psa;

extern "C" size_t psa_sa_conv1_prebn_workspace_bytes(int b, int n, int m, int c, int C1, int want_stats) {
    size_t bytes = 0;
    if (c > 0) bytes += ((size_t)b * n * C1 * sizeof(float) + 255) & ~(size_t)255;
    if (want_stats) {
        int q; dim3 g;
        f1_grid(b, m, &q, &g);
        bytes += (size_t)g.x * g.y * 2 * C1 * sizeof(float);
    }
    return bytes;
}

extern "C" int psa_sa_conv1_prebn(int b, int n, int m, int c, float radius, int nsample, const float* xyz,
                                  const float* new_xyz, const float* points, const float* w1, const float* bias, int C1,
                                  float* pre, int* idx, int* pts_cnt, float* stats, void* workspace,
                                  size_t workspace_bytes, psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 1 && m >= 0 && c >= 0 && nsample >= 1, "sa_conv1_prebn: bad dims b=%d n=%d m=%d c=%d nsample=%d", b, n, m, c, nsample);
    PSA_REQUIRE(C1 == 64 || C1 == 128, "sa_conv1_prebn: C1=%d must be 64 or 128", C1);
    if (b == 0 || m == 0) return PSA_OK;
    PSA_REQUIRE(xyz && new_xyz && w1 && pre && idx && (points || c == 0), "sa_conv1_prebn: null buffer");
    const size_t need = psa_sa_conv1_prebn_workspace_bytes(b, n, m, c, C1, stats != nullptr);
    PSA_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need), "sa_conv1_prebn: workspace of %zu bytes required", need);
    cudaStream_t st = as_stream(stream);
    F1Args a;
    a.n = n; a.m = m; a.nsample = nsample; a.C1 = C1;
    bool none = false;
    a.thr = ball_query_threshold(radius, &none);
    a.none = none ? 1 : 0;
    a.radius = radius;
    a.want_grid = (bq_grid_fits(n) && n >= 256 && m >= 32) ? 1 : 0;
    a.xyz = xyz; a.new_xyz = new_xyz; a.w1 = w1; a.bias = bias; a.pre = pre; a.idx = idx; a.pts_cnt = pts_cnt;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    a.uf = nullptr;
    if (c > 0) {
        int rc = launch_dense_raw((long long)b * n, c, C1, points, w1 + (size_t)3 * C1, reinterpret_cast<float*>(ws), st);
        if (rc != PSA_OK) return rc;
        a.uf = reinterpret_cast<float*>(ws);
        ws += ((size_t)b * n * C1 * sizeof(float) + 255) & ~(size_t)255;
    }
    dim3 grid;
    f1_grid(b, m, &a.q_per_cta, &grid);
    a.partial = stats ? reinterpret_cast<float*>(ws) : nullptr;
    size_t smem = bq_smem_bytes(n, a.want_grid != 0) + (size_t)kF1Warps * ((nsample + 3) & ~3) * sizeof(int) + (stats ? (size_t)kF1Warps * 2 * C1 * sizeof(float) : 0);
    PSA_SUPPORTED(smem <= 200 * 1024, "sa_conv1_prebn: n=%d exceeds the shared-memory resident limit", n);
#define PSA_F1_LAUNCH(NV_, ST_, PPT_)                                                                                        \
    do {                                                                                                                     \
        PSA_CUDA(cudaFuncSetAttribute(sa_conv1_prebn_kernel<NV_, ST_, PPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sa_conv1_prebn_kernel<NV_, ST_, PPT_><<<grid, kF1Warps * 32, smem, st>>>(a);                                         \
    } while (0)
#define PSA_F1_DISPATCH(ST_)                                                                                                 \
    do {                                                                                                                     \
        if (n <= 8 * kBqThreads) { if (C1 == 64) PSA_F1_LAUNCH(2, ST_, 8); else PSA_F1_LAUNCH(4, ST_, 8); }                  \
        else { if (C1 == 64) PSA_F1_LAUNCH(2, ST_, 16); else PSA_F1_LAUNCH(4, ST_, 16); }                                    \
    } while (0)
    // shared memory: ball-query arrays, per-warp idx rows, per-warp statistics (16-byte aligned: nsample rows of ints)
    if (stats) {
        PSA_F1_DISPATCH(true);
        f1_stats_reduce_kernel<<<(2 * C1 + 127) / 128, 128, 0, st>>>((int)(grid.x * grid.y), 2 * C1, a.partial, stats);
    } else {
        PSA_F1_DISPATCH(false);
    }
#undef PSA_F1_DISPATCH
#undef PSA_F1_LAUNCH
    return check_launch("sa_conv1_prebn_kernel");
}

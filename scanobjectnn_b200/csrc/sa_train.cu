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
#include "common.cuh"

namespace psa {

float ball_query_threshold(float radius, bool* none);   // grouping.cu

constexpr int kF1Warps = 8;

__device__ __forceinline__ float2 dist2_pair_f1(float2 x, float2 y, float2 z, float2 nqx, float2 nqy, float2 nqz) {
    const float2 dx = __fadd2_rn(x, nqx), dy = __fadd2_rn(y, nqy), dz = __fadd2_rn(z, nqz);
    float2 t = __fmul2_rn(dy, dy);
    t = __ffma2_rn(dx, dx, t);
    t = __ffma2_rn(dz, dz, t);
    return t;
}

// same scan as grouping.cu:ball_query_warp, but the idx row goes to shared memory (the conv stage reads it back)
__device__ __forceinline__ int ball_query_warp_smem(int n, int nsample, float thr, bool none, const float* sx, const float* sy,
                                                    const float* sz, float qx, float qy, float qz, int* idxrow, int lane) {
    int cnt = 0, first = -1;
    if (!none) {
        const float2 nqx = make_float2(-qx, -qx), nqy = make_float2(-qy, -qy), nqz = make_float2(-qz, -qz);
        const unsigned lt = lanemask_lt();
        for (int base = 0; base < n && cnt < nsample; base += 128) {
            const int k = base + lane * 4;
            const float4 X = *reinterpret_cast<const float4*>(sx + k);
            const float4 Y = *reinterpret_cast<const float4*>(sy + k);
            const float4 Z = *reinterpret_cast<const float4*>(sz + k);
            const float2 d01 = dist2_pair_f1(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), nqx, nqy, nqz);
            const float2 d23 = dist2_pair_f1(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), nqx, nqy, nqz);
            bool i0 = !(d01.x > thr), i1 = !(d01.y > thr), i2 = !(d23.x > thr), i3 = !(d23.y > thr);
            if (base + 128 > n) { i0 = i0 && (k < n); i1 = i1 && (k + 1 < n); i2 = i2 && (k + 2 < n); i3 = i3 && (k + 3 < n); }
            const unsigned m4 = (i0 ? 1u : 0u) | (i1 ? 2u : 0u) | (i2 ? 4u : 0u) | (i3 ? 8u : 0u);
            const unsigned anyb = __ballot_sync(0xffffffffu, m4 != 0u);
            if (anyb == 0u) continue;
            const unsigned b0 = __ballot_sync(0xffffffffu, i0), b1 = __ballot_sync(0xffffffffu, i1);
            const unsigned b2 = __ballot_sync(0xffffffffu, i2), b3 = __ballot_sync(0xffffffffu, i3);
            if (first < 0) {
                const int lf = __ffs(anyb) - 1;
                const unsigned mf = __shfl_sync(0xffffffffu, m4, lf);
                first = base + lf * 4 + (__ffs(mf) - 1);
            }
            int pos = cnt + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
            if (i0) { if (pos < nsample) idxrow[pos] = k; ++pos; }
            if (i1) { if (pos < nsample) idxrow[pos] = k + 1; ++pos; }
            if (i2) { if (pos < nsample) idxrow[pos] = k + 2; ++pos; }
            if (i3) { if (pos < nsample) idxrow[pos] = k + 3; }
            cnt += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
        }
    }
    if (cnt > nsample) cnt = nsample;
    const int fillv = first < 0 ? 0 : first;
    for (int l = cnt + lane; l < nsample; l += 32) idxrow[l] = fillv;
    __syncwarp();
    return cnt;
}

struct F1Args {
    int n, m, nsample, C1, q_per_cta;
    float thr;
    int none;
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

// lane owns channel pairs (2*lane + 64*i, 2*lane + 64*i + 1), i < NP = C1/64
template <int NP, bool STATS>
__global__ void __launch_bounds__(kF1Warps * 32)
sa_conv1_prebn_kernel(const __grid_constant__ F1Args a) {
    extern __shared__ __align__(16) float smem_f[];
    const int n = a.n, np = (n + 127) & ~127;
    float* sx = smem_f;
    float* sy = sx + np;
    float* sz = sy + np;
    int* srow = reinterpret_cast<int*>(sz + np);                 // kF1Warps * nsample
    float* sstat = reinterpret_cast<float*>(srow + kF1Warps * a.nsample);   // kF1Warps * 2 * C1 (STATS)
    const int cloud = blockIdx.y;
    const float* p1 = a.xyz + (size_t)cloud * n * 3;
    {
        const int total = n * 3;
        int i = threadIdx.x;
        for (; i + 7 * (kF1Warps * 32) < total; i += 8 * (kF1Warps * 32)) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(p1 + i + u * (kF1Warps * 32));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = i + u * (kF1Warps * 32), k = e / 3, c = e - k * 3;
                (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v[u];
            }
        }
        for (; i < total; i += kF1Warps * 32) {
            const int k = i / 3, c = i - k * 3;
            (c == 0 ? sx : (c == 1 ? sy : sz))[k] = __ldg(p1 + i);
        }
    }
    const float inf = __int_as_float(0x7f800000);
    for (int k = n + threadIdx.x; k < np; k += blockDim.x) { sx[k] = inf; sy[k] = inf; sz[k] = inf; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this lane's weight columns and bias
    float2 wx[NP], wy[NP], wz[NP], bs[NP], ssum[NP], ssq[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int c = 2 * lane + 64 * i;
        wx[i] = make_float2(__ldg(a.w1 + c), __ldg(a.w1 + c + 1));
        wy[i] = make_float2(__ldg(a.w1 + a.C1 + c), __ldg(a.w1 + a.C1 + c + 1));
        wz[i] = make_float2(__ldg(a.w1 + 2 * a.C1 + c), __ldg(a.w1 + 2 * a.C1 + c + 1));
        bs[i] = a.bias ? make_float2(__ldg(a.bias + c), __ldg(a.bias + c + 1)) : make_float2(0.f, 0.f);
        ssum[i] = make_float2(0.f, 0.f);
        ssq[i] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int q0 = blockIdx.x * a.q_per_cta;
    const int q1 = min(a.m, q0 + a.q_per_cta);
    const float* p2 = a.new_xyz + (size_t)cloud * a.m * 3;
    int* row = srow + warp * a.nsample;
    for (int q = q0 + warp; q < q1; q += kF1Warps) {
        const float qx = __ldg(p2 + q * 3 + 0), qy = __ldg(p2 + q * 3 + 1), qz = __ldg(p2 + q * 3 + 2);
        const int cnt = ball_query_warp_smem(n, a.nsample, a.thr, a.none != 0, sx, sy, sz, qx, qy, qz, row, lane);
        const size_t gq = (size_t)cloud * a.m + q;
        for (int l = lane; l < a.nsample; l += 32) a.idx[gq * a.nsample + l] = row[l];
        if (a.pts_cnt != nullptr && lane == 0) a.pts_cnt[gq] = cnt;
        float* outq = a.pre + gq * a.nsample * a.C1;
        for (int r = 0; r < a.nsample; ++r) {
            const int j = row[r];                                  // broadcast
            const float dx = sx[j] - qx, dy = sy[j] - qy, dz = sz[j] - qz;     // grouped_xyz - new_xyz (pointnet_util.py:46)
            const float* urow = a.uf ? a.uf + ((size_t)cloud * n + j) * a.C1 : nullptr;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                float2 u = make_float2(0.f, 0.f);
                if (urow) u = __ldg(reinterpret_cast<const float2*>(urow + 2 * lane + 64 * i));
                float2 v;
                v.x = fmaf(dz, wz[i].x, fmaf(dy, wy[i].x, fmaf(dx, wx[i].x, u.x))) + bs[i].x;
                v.y = fmaf(dz, wz[i].y, fmaf(dy, wy[i].y, fmaf(dx, wx[i].y, u.y))) + bs[i].y;
                __stcs(reinterpret_cast<float2*>(outq + (size_t)r * a.C1 + 2 * lane + 64 * i), v);   // streaming store
                if (STATS) {
                    ssum[i].x += v.x; ssum[i].y += v.y;
                    ssq[i].x = fmaf(v.x, v.x, ssq[i].x); ssq[i].y = fmaf(v.y, v.y, ssq[i].y);
                }
            }
        }
        __syncwarp();
    }
    if (STATS) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int c = 2 * lane + 64 * i;
            float* w = sstat + (size_t)warp * 2 * a.C1;
            w[c] = ssum[i].x; w[c + 1] = ssum[i].y;
            w[a.C1 + c] = ssq[i].x; w[a.C1 + c + 1] = ssq[i].y;
        }
        __syncthreads();
        float* dst = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * a.C1;
        for (int e = threadIdx.x; e < 2 * a.C1; e += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kF1Warps; ++w) s += sstat[(size_t)w * 2 * a.C1 + e];   // fixed order: deterministic
            dst[e] = s;
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
    PSA_REQUIRE(C1 >= 64 && C1 % 64 == 0 && C1 <= 256, "sa_conv1_prebn: C1=%d must be 64, 128, 192 or 256", C1);
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
    const int np = (n + 127) & ~127;
    size_t smem = (size_t)np * 3 * sizeof(float) + (size_t)kF1Warps * nsample * sizeof(int) + (stats ? (size_t)kF1Warps * 2 * C1 * sizeof(float) : 0);
    PSA_SUPPORTED(smem <= 200 * 1024, "sa_conv1_prebn: n=%d exceeds the shared-memory resident limit", n);
#define PSA_F1_LAUNCH(NP_, ST_)                                                                                              \
    do {                                                                                                                     \
        PSA_CUDA(cudaFuncSetAttribute(sa_conv1_prebn_kernel<NP_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sa_conv1_prebn_kernel<NP_, ST_><<<grid, kF1Warps * 32, smem, st>>>(a);                                               \
    } while (0)
    const int NP = C1 / 64;
    if (stats) {
        if (NP == 1) PSA_F1_LAUNCH(1, true); else if (NP == 2) PSA_F1_LAUNCH(2, true); else if (NP == 3) PSA_F1_LAUNCH(3, true); else PSA_F1_LAUNCH(4, true);
        f1_stats_reduce_kernel<<<(2 * C1 + 127) / 128, 128, 0, st>>>((int)(grid.x * grid.y), 2 * C1, a.partial, stats);
    } else {
        if (NP == 1) PSA_F1_LAUNCH(1, false); else if (NP == 2) PSA_F1_LAUNCH(2, false); else if (NP == 3) PSA_F1_LAUNCH(3, false); else PSA_F1_LAUNCH(4, false);
    }
#undef PSA_F1_LAUNCH
    return check_launch("sa_conv1_prebn_kernel");
}

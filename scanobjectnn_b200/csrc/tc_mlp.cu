// tc_mlp.cu -- the grouped shared MLP of a set-abstraction level on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// What the reference does (pointnet2/utils/pointnet_util.py:113-127): group_point -> (B,m,K,3+C) tensor -> three
// cuDNN 1x1 convs over B*m*K rows -> reduce_max.  What this kernel does per 128-row tile (128/K neighbourhoods):
//
//   layer 1   is never a GEMM over grouped rows.  (x_j - c) . Wx + f_j . Wf  =  U[j] + (x_j - c) . Wx   with
//             U = points . W1[3:,:] computed ONCE per source point (K-fold fewer rows, psa_sa_module_infer does it
//             with the dense kernel); each row-thread gathers its U row (512 B), adds the 3-term xyz part in FMAs,
//             applies the folded BN affine + ReLU, and writes the result straight into TENSOR MEMORY as the A operand
//             of layer 2 -- the (B,m,K,C) tensors of the reference never exist, not even in shared memory.
//   layers 2+ tcgen05.mma, A from TMEM (lane = row), B = weights RESIDENT in shared memory for the whole persistent
//             CTA (canonical K-major SWIZZLE_128B layout built once by the CTA), D in TMEM.  Between layers the
//             four row-warps pull D with tcgen05.ld, apply affine + ReLU, and push the next A operand with tcgen05.st.
//   max-pool  the last epilogue reduces each neighbourhood's rows with a transposing warp butterfly (31 shuffles per
//             32 columns) and writes (B,m,C_out) coalesced.
//
// fp32 parity on tf32/bf16 tensor cores: every product a*w is evaluated as three exactly-representable pieces
//   trunc_tf32(a)*trunc_tf32(w) + tf32(a - trunc(a))*trunc_tf32(w) + bf16(a)*bf16(w - trunc(w))
// (operands quantised by this code, so the tensor core sees exact values; fp32 accumulation in TMEM).  Operand-split
// error ~3e-6 relative (tests/test_mlp_gpu.py holds the whole chain to 1e-5 against fp64).  The bf16 third term keeps
// the resident weights at 6 B/element so that PointNet++'s 128->128->256 level fits in 227 KB when the last layer's
// output channels are split over two CTAs.
#include <float.h>

#include "common.cuh"
#include "mlp_internal.cuh"
#include "tc_common.cuh"

namespace psa {

using namespace tc;

constexpr int kTcThreads = 512;   // 16 warps: warp w works on TMEM lanes 32*(w%4).. (hardware rule) and column chunks w/4, w/4+4, ..
constexpr int kTcChunkWarps = kTcThreads / 128;
constexpr int kTmemCols = 512;
constexpr uint32_t D_COL = 0, AHI_COL = 128, ALO_COL = 256, ABF_COL = 384;
constexpr int kMaxTcLayers = 2;

struct TcArgs {
    long long groups;      // neighbourhoods = b*m
    int K;                 // rows per neighbourhood (32 | 64 | 128)
    int n, m;              // dataset points / queries per cloud
    const float* xyz;      // (b,n,3)
    const float* new_xyz;  // (b,m,3)
    const float* uf;       // (b*n, C1) = points . W1[3:,:], or null when the level has no input features
    const int* idx;        // (groups, K)
    float* out;            // (groups, Ntot[last])
    // layer 1 (FMA path)
    const float* w1x;      // (3, C1): rows 0..2 of W1
    const float* s1;       // scale or null
    const float* t1;       // shift
    int C1, relu1;
    // tensor layers
    int nl;
    const float* W[kMaxTcLayers];
    const float* s[kMaxTcLayers];
    const float* t[kMaxTcLayers];
    int relu[kMaxTcLayers];
    int Kd[kMaxTcLayers], Ntot[kMaxTcLayers];
    int nsplit;            // CTAs sharing one tile, each owning Ntot[last]/nsplit output channels
};

// quantise one row-chunk of 32 activations into the three A operands and store them into TMEM
__device__ __forceinline__ void store_a_chunk(uint32_t row_taddr, int ch, const float (&h)[32]) {
    uint32_t v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(tf32_trunc(h[q]));
    tmem_st32(row_taddr + AHI_COL + ch * 32, v);
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(tf32_trunc(h[q] - tf32_trunc(h[q])));
    tmem_st32(row_taddr + ALO_COL + ch * 32, v);
    uint32_t p[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = pack_bf16x2(h[2 * q], h[2 * q + 1]);
    tmem_st16(row_taddr + ABF_COL + ch * 16, p);
}

// Build the resident B operand of one tensor layer: W (global, [Kd][Ntot] row-major) columns [n_off, n_off+N) ->
// whi: trunc_tf32(w) as [N][Kd] K-major SW128 fp32;  wlo: bf16(w - trunc_tf32(w)) as [N][Kd] K-major SW128 bf16.
__device__ __forceinline__ void stage_weights(uint8_t* whi, uint8_t* wlo, const float* __restrict__ W, int Kd, int Ntot,
                                              int n_off, int N, int tid, int nthreads) {
    for (int e = tid; e < N * Kd; e += nthreads) {
        const int nn = e % N, k = e / N;
        const float w = __ldg(W + (size_t)k * Ntot + n_off + nn);
        const float hi = tf32_trunc(w);
        *reinterpret_cast<float*>(whi + swz_off_f32(nn, k, N)) = hi;
        *reinterpret_cast<__nv_bfloat16*>(wlo + swz_off_bf16(nn, k, N)) = __float2bfloat16_rn(w - hi);
    }
}

// one elected thread: D[128 x N] = A . W^T as the three-term split
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, uint32_t whi_addr, uint32_t wlo_addr, int Kd, int N) {
    const uint32_t d = tmem_base + D_COL;
    const uint32_t id_tf32 = make_idesc(kFmtTF32, 128, N);
    const uint32_t id_bf16 = make_idesc(kFmtBF16, 128, N);
    const uint32_t blk = (uint32_t)N * 128u;
    // The tensor core truncates (rounds toward zero) every time it adds into D, so the error grows with the number of
    // accumulation steps taken while D is large: run the two small correction terms first, the main term last.
    for (int s = 0; s < Kd / 16; ++s)
        mma_bf16_ts(d, tmem_base + ABF_COL + s * 8, make_smem_desc_sw128(wlo_addr + (s >> 2) * blk + (s & 3) * 32), id_bf16, s > 0);
    for (int s = 0; s < Kd / 8; ++s)
        mma_tf32_ts(d, tmem_base + ALO_COL + s * 8, make_smem_desc_sw128(whi_addr + (s >> 2) * blk + (s & 3) * 32), id_tf32, 1);
    for (int s = 0; s < Kd / 8; ++s)
        mma_tf32_ts(d, tmem_base + AHI_COL + s * 8, make_smem_desc_sw128(whi_addr + (s >> 2) * blk + (s & 3) * 32), id_tf32, 1);
}

// transposing butterfly: v[q] = column q of this lane's row; afterwards v[0] on lane l = max over the warp's 32 rows of column l
__device__ __forceinline__ float warp_colmax_32x32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            const float recv = __shfl_xor_sync(0xffffffffu, send, half);
            v[i] = fmaxf(keep, recv);
        }
    }
    return v[0];
}

struct TcSmemLayout {
    uint32_t whi[kMaxTcLayers], wlo[kMaxTcLayers];   // byte offsets from the 1024-aligned base
    uint32_t vec;                                    // float region: w1x[3*C1] s1[C1] t1[C1] then per layer s[N] t[N]
    uint32_t total;
};

__host__ __device__ inline TcSmemLayout tc_layout(const TcArgs& a) {
    TcSmemLayout L;
    uint32_t off = 0;
    for (int l = 0; l < a.nl; ++l) {
        const int N = (l == a.nl - 1) ? a.Ntot[l] / a.nsplit : a.Ntot[l];
        L.whi[l] = off; off += (uint32_t)N * a.Kd[l] * 4u;
        L.wlo[l] = off; off += (uint32_t)N * a.Kd[l] * 2u;
    }
    L.vec = off;
    off += 5u * a.C1 * 4u;
    for (int l = 0; l < a.nl; ++l) off += 2u * ((l == a.nl - 1) ? a.Ntot[l] / a.nsplit : a.Ntot[l]) * 4u;
    L.total = off;
    return L;
}

__global__ void __launch_bounds__(kTcThreads, 1)
tc_sa_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[kTcThreads / 32][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quarter = warp & 3, cs = warp >> 2;      // rows 32*quarter.., column-chunk slot
    const int row = quarter * 32 + lane;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const TcSmemLayout L = tc_layout(a);
    const int split = blockIdx.x % a.nsplit;
    const int worker = blockIdx.x / a.nsplit, nworkers = gridDim.x / a.nsplit;
    const int last = a.nl - 1;
    const int Nlast = a.Ntot[last] / a.nsplit;

    // ---- one-time setup: TMEM, barrier, resident weights, per-channel vectors ----
    if (warp == 0) tmem_alloc(&s_tmem, kTmemCols);
    if (tid == 0) { mbar_init(&s_mbar, 1); fence_mbar_init(); }
    float* vec = reinterpret_cast<float*>(base + L.vec);
    float* w1x = vec;                       // 3*C1
    float* s1 = vec + 3 * a.C1;
    float* t1 = s1 + a.C1;
    float* sl[kMaxTcLayers];
    float* tl[kMaxTcLayers];
    {
        float* p = t1 + a.C1;
        for (int l = 0; l < a.nl; ++l) {
            const int N = (l == last) ? Nlast : a.Ntot[l];
            sl[l] = p; tl[l] = p + N; p += 2 * N;
        }
    }
    for (int i = tid; i < 3 * a.C1; i += kTcThreads) w1x[i] = __ldg(a.w1x + i);
    for (int i = tid; i < a.C1; i += kTcThreads) { s1[i] = a.s1 ? __ldg(a.s1 + i) : 1.f; t1[i] = __ldg(a.t1 + i); }
    for (int l = 0; l < a.nl; ++l) {
        const int N = (l == last) ? Nlast : a.Ntot[l];
        const int n_off = (l == last) ? split * Nlast : 0;
        for (int i = tid; i < N; i += kTcThreads) {
            sl[l][i] = a.s[l] ? __ldg(a.s[l] + n_off + i) : 1.f;
            tl[l][i] = __ldg(a.t[l] + n_off + i);
        }
        stage_weights(base + L.whi[l], base + L.wlo[l], a.W[l], a.Kd[l], a.Ntot[l], n_off, N, tid, kTcThreads);
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = s_tmem;
    const uint32_t row_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t phase = 0;

    const int G = 128 / a.K;
    const long long ntiles = (a.groups + G - 1) / G;
    for (long long tile = worker; tile < ntiles; tile += nworkers) {
        const long long g0 = tile * G;
        const long long gid = g0 + row / a.K;
        const bool valid = gid < a.groups;
        // ---- layer 1 on the FMA pipe, straight into the A operand ----
        {
            float dx = 0.f, dy = 0.f, dz = 0.f;
            const float* urow = nullptr;
            if (valid) {
                const long long bi = gid / a.m;
                const int j = __ldg(a.idx + gid * a.K + (row % a.K));
                const float* p = a.xyz + ((size_t)bi * a.n + j) * 3;
                const float* c = a.new_xyz + (size_t)gid * 3;
                dx = __ldg(p) - __ldg(c); dy = __ldg(p + 1) - __ldg(c + 1); dz = __ldg(p + 2) - __ldg(c + 2);
                if (a.uf) urow = a.uf + ((size_t)bi * a.n + j) * a.C1;
            }
            for (int ch = cs; ch < a.C1 / 32; ch += kTcChunkWarps) {
                float h[32];
                if (urow) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(urow + ch * 32) + q);
                        h[4 * q] = u.x; h[4 * q + 1] = u.y; h[4 * q + 2] = u.z; h[4 * q + 3] = u.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) h[q] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int c = ch * 32 + q;
                    float pre = fmaf(dz, w1x[2 * a.C1 + c], fmaf(dy, w1x[a.C1 + c], fmaf(dx, w1x[c], h[q])));
                    float v = fmaf(pre, s1[c], t1[c]);
                    if (a.relu1) v = fmaxf(v, 0.f);
                    h[q] = valid ? v : 0.f;
                }
                store_a_chunk(row_taddr, ch, h);
            }
        }
        tmem_st_wait();
        fence_before_thread_sync();
        __syncthreads();
        for (int l = 0; l < a.nl; ++l) {
            const int N = (l == last) ? Nlast : a.Ntot[l];
            if (tid == 0) {
                fence_after_thread_sync();
                issue_layer(tmem_base, smem_u32(base + L.whi[l]), smem_u32(base + L.wlo[l]), a.Kd[l], N);
                mma_commit(&s_mbar);
            }
            mbar_wait(&s_mbar, phase);
            phase ^= 1u;
            fence_after_thread_sync();
            if (l != last) {
                for (int ch = cs; ch < N / 32; ch += kTcChunkWarps) {
                    uint32_t d[32];
                    tmem_ld32(row_taddr + D_COL + ch * 32, d);
                    tmem_ld_wait();
                    float h[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        float v = fmaf(__uint_as_float(d[q]), sl[l][ch * 32 + q], tl[l][ch * 32 + q]);
                        if (a.relu[l]) v = fmaxf(v, 0.f);
                        h[q] = valid ? v : 0.f;
                    }
                    store_a_chunk(row_taddr, ch, h);
                }
                tmem_st_wait();
                fence_before_thread_sync();
                __syncthreads();
            } else {
                // rows of one neighbourhood span K/32 lane-quarters; warps sharing a chunk slot combine through s_red
                const int quarters_per_group = a.K / 32;        // 1, 2 or 4
                const long long wg = g0 + (quarter * 32) / a.K; // this warp's neighbourhood
                const int nch = N / 32;
                for (int ch0 = 0; ch0 < nch; ch0 += kTcChunkWarps) {      // uniform trip count: barriers inside
                    const int ch = ch0 + cs;
                    float mx = -FLT_MAX;
                    if (ch < nch) {
                        uint32_t d[32];
                        tmem_ld32(row_taddr + D_COL + ch * 32, d);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            float x = fmaf(__uint_as_float(d[q]), sl[l][ch * 32 + q], tl[l][ch * 32 + q]);
                            if (a.relu[l]) x = fmaxf(x, 0.f);
                            v[q] = valid ? x : -FLT_MAX;
                        }
                        mx = warp_colmax_32x32(v, lane);
                    }
                    if (quarters_per_group > 1) {
                        s_red[warp][lane] = mx;
                        __syncthreads();
                        if ((quarter % quarters_per_group) == 0)
                            for (int o = 1; o < quarters_per_group; ++o) mx = fmaxf(mx, s_red[warp + o][lane]);
                        __syncthreads();
                    }
                    if (ch < nch && (quarter % quarters_per_group) == 0 && wg < a.groups)
                        a.out[(size_t)wg * a.Ntot[l] + split * Nlast + ch * 32 + lane] = mx;
                }
                fence_before_thread_sync();   // D fully read before the next tile's MMAs may overwrite it
            }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- self-test of one tensor layer: D[128 x N] = A[128 x K] . W[K x N] through exactly the device code above ----
__global__ void __launch_bounds__(kTcThreads, 1)
tc_selftest_kernel(int Kd, int N, const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3, cs = warp >> 2;
    const int row = quarter * 32 + lane;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* whi = base;
    uint8_t* wlo = base + (size_t)N * Kd * 4;
    if (warp == 0) tmem_alloc(&s_tmem, kTmemCols);
    if (tid == 0) { mbar_init(&s_mbar, 1); fence_mbar_init(); }
    stage_weights(whi, wlo, W, Kd, N, 0, N, tid, kTcThreads);
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = s_tmem;
    const uint32_t row_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int ch = cs; ch < Kd / 32; ch += kTcChunkWarps) {
        float h[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) h[q] = A[(size_t)row * Kd + ch * 32 + q];
        store_a_chunk(row_taddr, ch, h);
    }
    tmem_st_wait();
    fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_thread_sync();
        issue_layer(tmem_base, smem_u32(whi), smem_u32(wlo), Kd, N);
        mma_commit(&s_mbar);
    }
    mbar_wait(&s_mbar, 0);
    fence_after_thread_sync();
    for (int ch = cs; ch < N / 32; ch += kTcChunkWarps) {
        uint32_t d[32];
        tmem_ld32(row_taddr + D_COL + ch * 32, d);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) D[(size_t)row * N + ch * 32 + q] = __uint_as_float(d[q]);
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// Can this MLP / geometry run on the tensor-core kernel?  (otherwise the fp32-FMA fused kernel in mlp.cu is used)
bool tc_sa_eligible(const psa_mlp* mlp, int c, int nsample, TcArgs* out) {
    if (mlp->n_layers < 2 || mlp->n_layers > 1 + kMaxTcLayers) return false;
    if (!(nsample == 32 || nsample == 64 || nsample == 128)) return false;
    if (mlp->channels[0] != 3 + c) return false;
    const int C1 = mlp->channels[1];
    if (!(C1 == 64 || C1 == 128)) return false;
    TcArgs a{};
    a.C1 = C1;
    a.nl = mlp->n_layers - 1;
    for (int l = 0; l < a.nl; ++l) {
        a.Kd[l] = mlp->channels[1 + l];
        a.Ntot[l] = mlp->channels[2 + l];
        if (!(a.Kd[l] == 64 || a.Kd[l] == 128)) return false;
        const bool is_last = (l == a.nl - 1);
        if (!is_last && !(a.Ntot[l] == 64 || a.Ntot[l] == 128)) return false;
        if (is_last && !(a.Ntot[l] == 64 || a.Ntot[l] % 128 == 0)) return false;
    }
    a.nsplit = a.Ntot[a.nl - 1] <= 128 ? 1 : a.Ntot[a.nl - 1] / 128;
    if (a.nsplit > 4) return false;
    if (tc_layout(a).total + 1024 > 220 * 1024) return false;
    *out = a;
    return true;
}

int launch_tc_sa(TcArgs& a, cudaStream_t st) {
    const TcSmemLayout L = tc_layout(a);
    size_t smem = (size_t)L.total + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: each CTA allocates all 512 TMEM columns
    PSA_CUDA(cudaFuncSetAttribute(tc_sa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int G = 128 / a.K;
    const long long ntiles = (a.groups + G - 1) / G;
    long long workers = kNumSMs / a.nsplit;
    if (workers > ntiles) workers = ntiles;
    if (workers < 1) workers = 1;
    tc_sa_kernel<<<(int)(workers * a.nsplit), kTcThreads, smem, st>>>(a);
    return check_launch("tc_sa_kernel");
}

}  // namespace psa

using namespace psa;

// Diagnostic entry point (not part of the reference's op surface): one 128-row tile through one tensor-core layer.
extern "C" PSA_API int psa_tc_selftest(int Kd, int N, const float* A, const float* W, float* D, psa_stream_t stream) {
    PSA_REQUIRE((Kd == 64 || Kd == 128) && (N == 64 || N == 128), "tc_selftest: Kd, N must be 64 or 128");
    size_t smem = (size_t)N * Kd * 6 + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;
    PSA_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, kTcThreads, smem, as_stream(stream)>>>(Kd, N, A, W, D);
    return check_launch("tc_selftest_kernel");
}

// 0 = auto (tensor cores where the shapes allow, fp32 FMA otherwise); 1 = always the fp32-FMA kernels
static int g_mlp_mode = 0;
extern "C" PSA_API int psa_set_mlp_mode(int mode) {
    PSA_REQUIRE(mode == 0 || mode == 1, "set_mlp_mode: mode must be 0 (auto) or 1 (fp32 FMA)");
    g_mlp_mode = mode;
    return PSA_OK;
}
extern "C" PSA_API int psa_get_mlp_mode(void) { return g_mlp_mode; }

extern "C" size_t psa_sa_module_workspace_bytes(int b, int n, int m, int c, int nsample, const psa_mlp* mlp) {
    (void)m;
    TcArgs a;
    if (g_mlp_mode == 0 && mlp != nullptr && c > 0 && tc_sa_eligible(mlp, c, nsample, &a))
        return (size_t)b * n * a.C1 * sizeof(float);
    return 0;
}

extern "C" int psa_sa_module_infer(int b, int n, int m, int c, float radius, int nsample, const float* xyz,
                                   const float* new_xyz, const float* points, const int* idx_in, const psa_mlp* mlp,
                                   float* out, int* idx_out, int* pts_cnt, void* workspace, size_t workspace_bytes,
                                   psa_stream_t stream) {
    int rc = validate_mlp_public(mlp, "sa_module");
    if (rc != PSA_OK) return rc;
    PSA_REQUIRE(b >= 0 && n >= 1 && m >= 0 && c >= 0 && nsample >= 1, "sa_module: bad dims b=%d n=%d m=%d c=%d nsample=%d", b, n, m, c, nsample);
    PSA_REQUIRE(mlp->channels[0] == 3 + c, "sa_module: mlp input width %d != 3 + c (%d)", mlp->channels[0], 3 + c);
    if (b == 0 || m == 0) return PSA_OK;
    PSA_REQUIRE(xyz && new_xyz && out && (points || c == 0), "sa_module: null buffer");
    const int* idx = idx_in;
    if (idx == nullptr) {
        PSA_REQUIRE(idx_out != nullptr, "sa_module: idx_out must be provided when idx_in is NULL (it receives the ball query)");
        rc = psa_query_ball_point(b, n, m, radius, nsample, xyz, new_xyz, idx_out, pts_cnt, stream);
        if (rc != PSA_OK) return rc;
        idx = idx_out;
    }
    cudaStream_t st = as_stream(stream);
    TcArgs a;
    if (g_mlp_mode == 0 && tc_sa_eligible(mlp, c, nsample, &a)) {
        a.groups = (long long)b * m; a.K = nsample; a.n = n; a.m = m;
        a.xyz = xyz; a.new_xyz = new_xyz; a.idx = idx; a.out = out; a.uf = nullptr;
        a.w1x = mlp->weight[0]; a.s1 = mlp->scale[0]; a.t1 = mlp->shift[0]; a.relu1 = mlp->relu[0];
        for (int l = 0; l < a.nl; ++l) {
            a.W[l] = mlp->weight[1 + l]; a.s[l] = mlp->scale[1 + l]; a.t[l] = mlp->shift[1 + l]; a.relu[l] = mlp->relu[1 + l];
        }
        if (c > 0) {
            const size_t need = (size_t)b * n * a.C1 * sizeof(float);
            PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need,
                        "sa_module: workspace of %zu bytes required (psa_sa_module_workspace_bytes), got %zu", need, workspace_bytes);
            // U = points . W1[3:,:]  once per source point (rows b*n), raw (affine + ReLU are applied after the xyz part)
            DenseArgs d;
            d.rows = (long long)b * n; d.K = c; d.N = a.C1; d.pool_k = 1; d.relu = 0;
            d.x = points; d.W = mlp->weight[0] + (size_t)3 * a.C1; d.scale = nullptr; d.shift = nullptr;
            d.out = reinterpret_cast<float*>(workspace);
            rc = launch_dense(d, st);
            if (rc != PSA_OK) return rc;
            a.uf = d.out;
        }
        return launch_tc_sa(a, st);
    }
    return sa_module_simt(b, n, m, c, nsample, xyz, new_xyz, points, idx, mlp, out, st);
}

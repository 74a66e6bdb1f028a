// tc_mlp.cu -- the grouped shared MLP of a set-abstraction level and the dense layers on the 5th-gen tensor cores
// (tcgen05 + TMEM), hand-written PTX wrappers in tc_common.cuh.
//
// What the reference does (pointnet2/utils/pointnet_util.py:113-127): group_point -> (B,m,K,3+C) tensor -> three
// cuDNN 1x1 convs over B*m*K rows -> reduce_max.  What the kernels here do per 128-row tile (128/K neighbourhoods):
//
//   layer 1   is never a GEMM over grouped rows.  (x_j - c) . Wx + f_j . Wf  =  U[j] + (x_j - c) . Wx   with
//             U = points . W1[3:,:] computed ONCE per source point (K-fold fewer rows, a dense-layer launch); each
//             row-thread gathers its U row (512 B), adds the 3-term xyz part in FMAs, applies the folded BN affine + ReLU
//             and writes the result straight into TENSOR MEMORY as the A operand of layer 2 -- the (B,m,K,C) tensors of
//             the reference never exist, not even in shared memory.
//   layers 2+ tcgen05.mma, A from TMEM (lane = row), B = weights in shared memory in the canonical K-major SWIZZLE_128B
//             layout, dropped there by cp.async.bulk from pre-arranged images; D in TMEM.  Between layers the row warps
//             pull D with tcgen05.ld, apply affine + ReLU and push the next A operand with tcgen05.st.
//   max-pool  the last epilogue reduces each neighbourhood's rows with a transposing warp butterfly and writes
//             (B,m,C_out) coalesced.  64-wide levels ending in a 128-wide layer (DB = 3, SA1) run the LAST layer transposed
//             instead: H^T written to shared memory by the previous epilogue, D^T[channel][row] -- lane = channel, so the
//             max-pool is an in-thread reduction, the affine a per-thread scalar and the stores coalesced row segments.
//
// Kernels (all templated on NP, the pieces per operand -- Split<NP> in tc_common.cuh):
//   tc_sa_dual_kernel<DB,NP>  SA level, two row groups per CTA, per-group streamed last layer, tensor-pipe token, one or two D
//                             slots (DB = 0/1/2), DB = 3 = transposed last layer; optional centre weights = multi-layer EdgeConv over 3-D points
//   tc_dense3_kernel<NP>      dense layer, transposed (lane = channel), both operands from shared memory; also the
//                             training-mode forward (previous batch norm applied on load, statistics in the epilogue)
//   tc_dense2_kernel<NT,NP>   dense layer, A from TMEM, warp-specialised pipeline (N = 64 or K > 512)
// fp32 parity: operands are quantised by this code, so the tensor core only ever sees exactly representable values; fp32
// accumulation in TMEM truncates, hence small terms first and K cut into <= 128-wide pieces (tests hold 1e-5 vs fp64).
//   NP = 2 (inference default): two fp16 pieces, three MMAs per product; every kernel tracks the leading pieces it stores and raises
//           a device-side flag when a value leaves the fp16 range -- the launchers then rerun the op on the NP = 3 instantiation
//           (enqueued unconditionally, a no-op unless the flag is set: `run_if`);
//   NP = 3 (psa_set_mlp_mode(2), the guarded rerun, the training forward): three bf16 pieces, six MMAs per product.
// Levels the dual kernel cannot hold run on the fp32-FMA fused kernel of mlp.cu.
#include <float.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "mlp_internal.cuh"
#include "tc_common.cuh"

namespace psa {

using namespace tc;

constexpr int kMaxTcLayers = 2;

// ------------------------------------------------------------------------------------------------------------------
// Weight images.  A tensor layer W (K x N, row-major, fp32) is pre-arranged once per weight set (tc_prep_weights_kernel) into
// blocks that can be dropped into shared memory by a single cp.async.bulk and fed to tcgen05.mma unchanged:
//   block (nt, kc) covers output channels [nt*Nt, nt*Nt+Nt) x input channels [kc*64, kc*64+64): NP 16-bit pieces
//   (every piece exactly representable), each [Nt][64] K-major SWIZZLE_128B, Nt*128 B per piece; blocks stored in (nt major,
//   kc minor) order.  2 NP bytes per weight.
// ------------------------------------------------------------------------------------------------------------------
// np = pieces per weight: 3 (bf16x3, 6 bytes per weight) or 2 (fp16x2, 4 bytes); see Split<NP> in tc_common.cuh
__host__ __device__ inline uint32_t tc_block_bytes(int Nt, int np) { return (uint32_t)Nt * 128u * (uint32_t)np; }
__host__ __device__ inline size_t tc_image_bytes(int K, int N, int np) { return (size_t)K * N * 2u * (size_t)np; }     // independent of the tile width
// an image allocation = the blocks + a 256-byte trailer whose first word is set when a weight left the fp16 range (np = 2)
__host__ __device__ inline size_t tc_image_alloc_bytes(int K, int N, int np) { return ((tc_image_bytes(K, N, np) + 255) & ~(size_t)255) + 256; }

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
    const float* w1c;      // optional (3, C1): weights applied to the CENTRE coordinates new_xyz (EdgeConv's x_i part), or null
    const float* w1x;      // (3, C1): rows 0..2 of W1
    const float* s1;       // scale or null
    const float* t1;       // shift
    int C1, relu1;
    // tensor layers
    int nl;
    const uint8_t* image[kMaxTcLayers];   // weight images (global)
    const float* s[kMaxTcLayers];
    const float* t[kMaxTcLayers];
    int relu[kMaxTcLayers];
    int Kd[kMaxTcLayers], Ntot[kMaxTcLayers];
    int stream_last;       // 1: the last layer's weights do not fit next to the others -> one 128-channel tile at a time
    int ntcap;             // 64 (narrow configuration) or 128 (wide)
    int dual;              // 1: tc_sa_dual_kernel (two row groups per CTA, 64-wide output tiles)
    unsigned int* tile_counter;   // zeroed before the launch: tiles are handed out dynamically (CTAs that start late or
                                  // share their SM with another stream's kernels simply take fewer)
    int np;                       // operand pieces: 2 (fp16x2) or 3 (bf16x3)
    int tmode;                    // 1: the last layer runs TRANSPOSED (tc_sa_dual_kernel<3, 2>, see dual_dcol): lane = channel
    unsigned int* ovf;            // np = 2: set to 1 when an activation or weight left the fp16 range (the result is then invalid)
    const unsigned int* run_if;   // non-null: the launch is a no-op unless *run_if != 0 (the np = 3 rerun of a flagged launch)
    const unsigned int* wflag[kMaxTcLayers];   // np = 2: trailer word of each weight image
};

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

#ifdef PSA_TC_TIMING
// debug build only (tools/tc_timing.py): cycles thread 0 of every CTA spends in each phase of the tensor-core kernels
__device__ unsigned long long g_tc_timing[8];
#define TC_STAMP(i) do { if (tid == 0) { const long long now_ = clock64(); tacc[i] += (unsigned long long)(now_ - tprev); tprev = now_; } } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------------------------
// tc_sa_dual_kernel -- levels with 128-wide layers (PointNet++ SA2: 131 -> 128 -> 128 -> 256 over 64-point neighbourhoods).
//
// One tile at a time per CTA serialises "row work" (gather, BN/ReLU, operand split, max-pool: ~56 % of a tile) and the MMAs
// (~44 %); two CTAs per SM would hide one behind the other, but a 128-wide level does not fit twice (TMEM columns, 96 KB of
// weights).  This kernel gets the overlap inside ONE CTA:
//   * two independent ROW GROUPS of 8 warps, each with its own 128-row tile, its own 256 TMEM columns, its own MMA
//     issuer (thread 0 of the group), mbarriers, named barrier and tile claims; while one group gathers / pools, the
//     other group's MMAs run;
//   * the weights of the inner layer are resident ONCE and shared by both groups; a last layer that does not fit is
//     streamed per group, one 64-channel tile (48 KB) at a time, L2 -> shared memory during the previous epilogue;
//   * operands are split into NP 16-bit pieces each (Split<NP>): the A operand of a K = 128 layer is 64 NP TMEM columns, so that
//     A + a 64-column D (two of them with NP = 2) fit in a group's 256; weights are 2 NP bytes per element.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kImageBf16x3 = 0x100;      // flags in psa_mlp.image_nt / psa_mlp_image_plan: image holds three bf16 pieces ..
constexpr int kImageF16x2 = 0x200;       // .. or two fp16 pieces
constexpr int kImageFlags = kImageBf16x3 | kImageF16x2;
__host__ __device__ inline int image_flag(int np) { return np == 2 ? kImageF16x2 : kImageBf16x3; }

template <int NP>
__global__ void tc_prep_weights_kernel(int K, int Kp, int N, int Nt, const float* __restrict__ W, uint8_t* __restrict__ image,
                                       unsigned int* __restrict__ trailer, const unsigned int* __restrict__ run_if) {
    if (run_if != nullptr && *run_if == 0u) return;
    const int KC = Kp / 64;
    const uint32_t bb = tc_block_bytes(Nt, NP), piece = (uint32_t)Nt * 128u;
    uint32_t ovf = 0u;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Kp * N; e += gridDim.x * blockDim.x) {
        const int n = e % N, k = e / N;
        const float w = k < K ? __ldg(W + e) : 0.f;
        uint32_t pc[NP];
        split_pair<NP>(w, 0.f, pc, ovf);
        uint8_t* blk = image + (size_t)((n / Nt) * KC + (k >> 6)) * bb;
        const uint32_t off = swz_off_bf16(n % Nt, k & 63, Nt);
#pragma unroll
        for (int i = 0; i < NP; ++i) *reinterpret_cast<uint16_t*>(blk + i * piece + off) = (uint16_t)(pc[i] & 0xffffu);
    }
    if (NP == 2 && f16x2_overflowed(ovf)) atomicOr(trailer, 1u);
}

struct TcDual {
    static constexpr int kThreads = 512, kGroupThreads = 256, kGroupCols = 256, kNt = 64;
    static constexpr uint32_t D = 0, A1 = 64, A2 = 128, A3 = 192;
};

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;\n" ::"r"(g + 1) : "memory"); }

// split 32 activations into NP pieces and store them as 16 columns each at a1, a1 + ps, .. (h is clobbered)
template <int NP>
__device__ __forceinline__ void store_a_at(uint32_t a1, float (&h)[32], uint32_t ps, uint32_t& ovf) {
    uint32_t p[16];
    if constexpr (NP == 3) {
        (void)ovf;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p[q] = pack_bf16x2(h[2 * q], h[2 * q + 1]);
            h[2 * q] -= __uint_as_float(p[q] << 16);
            h[2 * q + 1] -= __uint_as_float(p[q] & 0xffff0000u);
        }
        tmem_st16(a1, p);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p[q] = pack_bf16x2(h[2 * q], h[2 * q + 1]);
            h[2 * q] -= __uint_as_float(p[q] << 16);
            h[2 * q + 1] -= __uint_as_float(p[q] & 0xffff0000u);
        }
        tmem_st16(a1 + ps, p);
#pragma unroll
        for (int q = 0; q < 16; ++q) p[q] = pack_bf16x2(h[2 * q], h[2 * q + 1]);
        tmem_st16(a1 + 2 * ps, p);
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p[q] = pack_f16x2(h[2 * q], h[2 * q + 1]);
            track_f16x2(ovf, p[q]);
            const float2 f = unpack_f16x2(p[q]);
            h[2 * q] -= f.x;
            h[2 * q + 1] -= f.y;
        }
        tmem_st16(a1, p);
#pragma unroll
        for (int q = 0; q < 16; ++q) p[q] = pack_f16x2(h[2 * q], h[2 * q + 1]);
        tmem_st16(a1 + ps, p);
    }
}

// the same for activations held as 16 float2 (the dual kernel's row work runs on the packed FFMA2 / FADD2 pipe: two channels per
// instruction); the fp16 residual h - f32(a1) is one FFMA2 with (-1, -1)
template <int NP>
__device__ __forceinline__ void store_a2_at(uint32_t a1, float2 (&h)[16], uint32_t ps, uint32_t& ovf) {
    uint32_t p[16];
    if constexpr (NP == 3) {
        (void)ovf;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p[q] = pack_bf16x2(h[q].x, h[q].y);
            h[q].x -= __uint_as_float(p[q] << 16);
            h[q].y -= __uint_as_float(p[q] & 0xffff0000u);
        }
        tmem_st16(a1, p);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p[q] = pack_bf16x2(h[q].x, h[q].y);
            h[q].x -= __uint_as_float(p[q] << 16);
            h[q].y -= __uint_as_float(p[q] & 0xffff0000u);
        }
        tmem_st16(a1 + ps, p);
#pragma unroll
        for (int q = 0; q < 16; ++q) p[q] = pack_bf16x2(h[q].x, h[q].y);
        tmem_st16(a1 + 2 * ps, p);
    } else {
        const float2 neg1 = make_float2(-1.f, -1.f);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p[q] = pack_f16x2(h[q].x, h[q].y);
            track_f16x2(ovf, p[q]);
            h[q] = __ffma2_rn(unpack_f16x2(p[q]), neg1, h[q]);       // exact: a product with -1, then one rounding of the difference
        }
        tmem_st16(a1, p);
#pragma unroll
        for (int q = 0; q < 16; ++q) p[q] = pack_f16x2(h[q].x, h[q].y);
        tmem_st16(a1 + ps, p);
    }
}

// tmode: 32 activations of row `row` (input channels [k0, k0 + 32)) split into NP pieces and written as the B operand of the transposed
// last layer: [128 rows][64 k] K-major SWIZZLE_128B per piece (16 KB), four 16-byte chunks per piece (conflict-free per quarter warp)
template <int NP>
__device__ __forceinline__ void store_h_smem(uint8_t* hbase, int row, int k0, float2 (&h)[16], uint32_t& ovf) {
    static_assert(NP == 2, "the transposed mode is instantiated for fp16x2 operands");
    uint32_t p[16];
    const float2 neg1 = make_float2(-1.f, -1.f);
    const uint32_t rbase = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        p[q] = pack_f16x2(h[q].x, h[q].y);
        track_f16x2(ovf, p[q]);
        h[q] = __ffma2_rn(unpack_f16x2(p[q]), neg1, h[q]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(hbase + rbase + ((((uint32_t)(k0 >> 3) + j) ^ (uint32_t)(row & 7)) << 4)) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = pack_f16x2(h[q].x, h[q].y);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(hbase + 16384u + rbase + ((((uint32_t)(k0 >> 3) + j) ^ (uint32_t)(row & 7)) << 4)) =
            make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

// issuer warp (converged): D[128 x NT_] (+)= sum over the piece pairs of Split<NP>, KC blocks of 64 input channels.
// a1_col: TMEM column of piece 1 of the A operand (pieces PS columns apart); d_col: accumulator (overwritten by the first MMA).
template <int KC, int NT_, int PS, int NP>
__device__ __forceinline__ void issue_tile_c(uint32_t tmem_base, uint32_t d_col, uint32_t a1_col, uint32_t blocks_addr) {
    constexpr uint32_t bb = NT_ * 128u * NP, piece = NT_ * 128u;
    const uint32_t tb = warp_uniform(tmem_base);
    const uint32_t d = tb + d_col;
    const uint32_t a1 = tb + a1_col;
    const uint32_t idesc = make_idesc(Split<NP>::kFmt, 128, NT_);
    const SmemDescBase b0 = smem_desc_base(warp_uniform(blocks_addr));
#pragma unroll
    for (int t = 0; t < Split<NP>::kTerms; ++t)
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
                mma_bf16_ts(d, a1 + Split<NP>::a(t) * PS + kc * 32 + s4 * 8, smem_desc_at(b0, kc * bb + Split<NP>::w(t) * piece + s4 * 32), idesc,
                            (t | kc | s4) ? 1u : 0u);
}
// TMEM columns of a row group (256), by D-buffering mode DB:
//   0  one D slot:            D 0..63 | A pieces 64 columns apart from 64
//   1  levels whose layers are all <= 64 wide: D slots 0 and 64 | A pieces 32 columns apart from 128
//   2  two-piece operands (NP = 2) of 128-wide layers leave 192..255 free: D slots 0 and 192 | A pieces at 64 and 128
//   3  64-wide levels whose last layer is 128 wide, NP = 2: inner layer as in mode 1 (D 0..63, A pieces from 128); the LAST layer runs
//      transposed, D^T[128 channels][128 rows] at 0..127 = W^T (A operand, shared memory) x H^T (B operand, shared memory, written by
//      the previous epilogue) -- lane = channel: the max-pool is an in-thread reduction, the affine a per-thread scalar, the output
//      store a coalesced 128-byte row segment
template <int DB> __device__ __forceinline__ constexpr uint32_t dual_dcol(int dslot) { return dslot == 0 ? 0u : (DB == 2 ? 192u : 64u); }
template <int NP, int DB>
__device__ __forceinline__ void issue_tile(uint32_t gbase, uint32_t blocks_addr, int KC, int dslot) {
    if constexpr (DB == 1 || DB == 3) {
        if (dslot == 0) issue_tile_c<1, TcDual::kNt, 32, NP>(gbase, 0, 128, blocks_addr);
        else issue_tile_c<1, TcDual::kNt, 32, NP>(gbase, 64, 128, blocks_addr);
    } else if constexpr (DB == 2) {
        if (KC == 1) {
            if (dslot == 0) issue_tile_c<1, TcDual::kNt, 64, NP>(gbase, 0, TcDual::A1, blocks_addr);
            else issue_tile_c<1, TcDual::kNt, 64, NP>(gbase, 192, TcDual::A1, blocks_addr);
        } else {
            if (dslot == 0) issue_tile_c<2, TcDual::kNt, 64, NP>(gbase, 0, TcDual::A1, blocks_addr);
            else issue_tile_c<2, TcDual::kNt, 64, NP>(gbase, 192, TcDual::A1, blocks_addr);
        }
    } else {
        if (KC == 1) issue_tile_c<1, TcDual::kNt, 64, NP>(gbase, TcDual::D, TcDual::A1, blocks_addr);
        else issue_tile_c<2, TcDual::kNt, 64, NP>(gbase, TcDual::D, TcDual::A1, blocks_addr);
    }
}

struct TcDualLayout {
    uint32_t hbuf[2];            // tmode: per group, the last layer's B operand [128 rows][64 k] K-major SWIZZLE_128B, np pieces of 16 KB
    uint32_t w[kMaxTcLayers];    // resident layers
    uint32_t ring[2];            // per-group ring (one 64-channel tile of the streamed last layer)
    uint32_t ring_bytes;
    uint32_t vec, total;
};

__host__ __device__ inline TcDualLayout tc_dual_layout(const TcArgs& a) {
    TcDualLayout L;
    uint32_t off = 0;
    for (int l = 0; l < a.nl; ++l) {
        L.w[l] = off;
        if (!(a.stream_last && l == a.nl - 1)) off += (uint32_t)tc_image_bytes(a.Kd[l], a.Ntot[l], a.np);
    }
    L.ring_bytes = a.stream_last ? (uint32_t)(a.Kd[a.nl - 1] / 64) * tc_block_bytes(TcDual::kNt, a.np) : 0u;
    L.ring[0] = off; off += L.ring_bytes;
    L.ring[1] = off; off += L.ring_bytes;
    L.hbuf[0] = off; if (a.tmode) off += (uint32_t)a.np * 16384u;
    L.hbuf[1] = off; if (a.tmode) off += (uint32_t)a.np * 16384u;
    L.vec = off;
    off += 8u * a.C1 * 4u;       // w1x (3 C1), s1, t1, w1c (3 C1)
    for (int l = 0; l < a.nl; ++l) off += 2u * a.Ntot[l] * 4u;
    L.total = off;
    return L;
}

// D chunk (32 columns of this lane's row) -> relu?(d * scale + shift).  Rows past the end of the problem need no masking: a row of D
// depends on the same row of A only, validity is per NEIGHBOURHOOD (K divides the tile), and outputs of neighbourhoods past the end are
// never stored; their layer-1 inputs are zeros, so they carry finite values (no spurious range flag either).
__device__ __forceinline__ void affine_chunk(const uint32_t (&d)[32], const float* __restrict__ sc_, const float* __restrict__ sh_, int relu,
                                             float2 (&h)[16]) {
    const float4* s4 = reinterpret_cast<const float4*>(sc_);
    const float4* t4 = reinterpret_cast<const float4*>(sh_);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 sc = s4[q], sh = t4[q];
        float2 v0 = __ffma2_rn(make_float2(__uint_as_float(d[4 * q + 0]), __uint_as_float(d[4 * q + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
        float2 v1 = __ffma2_rn(make_float2(__uint_as_float(d[4 * q + 2]), __uint_as_float(d[4 * q + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
        if (relu) { v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f); }
        h[2 * q] = v0; h[2 * q + 1] = v1;
    }
}

// DBUF (levels whose layers are all <= 64 wide, nothing streamed): the A operand is 3 x 32 columns (at 128..), which leaves
// room for two D slots (0, 64) -- the last layer's tiles are issued in pairs and the second tile's MMAs run under the
// first tile's pooled epilogue.
// The two row groups of a CTA have identical phase lengths, so left alone they run in lock-step: both in row work
// (fighting for issue slots), then both in their MMA batch (each at half the tensor rate).  A token makes the batches
// run one after the other: the first group finishes at full rate and its row work then overlaps the second group's
// MMAs -- the groups fall into anti-phase.  Released right after the batch is ISSUED (the pipe executes in order).
// (branch-free at the source level -- predicated atomics + a shuffle -- so the issuer warp's control flow stays uniform)
__device__ __forceinline__ void pipe_acquire(int* token, int lane) {
    (void)lane;
    uint32_t busy;
    do {
        uint32_t r;
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "mov.b32 %0, 1;\n\t"
            "@q atom.shared.cas.b32 %0, [%1], 0, 1;\n\t}\n"
            : "=r"(r)
            : "r"(smem_u32(token))
            : "memory");
        busy = __shfl_sync(0xffffffffu, r, 0);
    } while (busy != 0u);
    __syncwarp();
}
__device__ __forceinline__ void pipe_release(int* token, int lane) {
    (void)lane;
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b32 t;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q atom.shared.exch.b32 t, [%0], 0;\n\t}\n" ::"r"(smem_u32(token))
        : "memory");
}

template <int DB, int NP>
__global__ void __launch_bounds__(TcDual::kThreads, 1)
tc_sa_dual_kernel(const __grid_constant__ TcArgs a) {
    constexpr bool DBUF = DB != 0;
    if (a.run_if != nullptr && *a.run_if == 0u) return;      // np = 3 rerun of a launch that stayed inside the fp16 range: nothing to do
    constexpr int kNt = TcDual::kNt;
    uint32_t ovf = 0u;                                        // np = 2: packed |max| of every leading piece this thread stores
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar[2];  // MMA completion, per group
    __shared__ __align__(8) uint64_t s_mbar2[2]; // MMA completion of the second D slot (dbuf levels), per group
    __shared__ __align__(8) uint64_t s_wbar[2];  // ring tile landed, per group
    __shared__ __align__(8) uint64_t s_rbar;     // resident weights landed
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[TcDual::kThreads / 32][32];
    __shared__ unsigned int s_tile[2][2];
    __shared__ __align__(16) float4 s_geo[2][128];   // per group: (dx, dy, dz, source row) of the NEXT tile's rows, prefetched
    __shared__ int s_token;                          // tensor-pipe token: the two groups' MMA batches run one after the other
    __shared__ int s_nonneg;                         // every scale of the last layer >= 0: pool first, affine + ReLU after

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_u = (int)warp_uniform((uint32_t)warp);
    const int g = warp_u >> 3, gtid = tid & 255;
    const bool issuer = (warp_u & 7) == 0;    // first warp of each group: converged MMA issue (tc_common.cuh), ring refills
    const int quarter = warp_u & 3, cs = (warp_u >> 2) & 1;
    const int row = quarter * 32 + lane;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const TcDualLayout L = tc_dual_layout(a);
    const int last = a.nl - 1;

    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1); mbar_init(&s_mbar2[0], 1); mbar_init(&s_mbar2[1], 1); mbar_init(&s_wbar[0], 1); mbar_init(&s_wbar[1], 1); mbar_init(&s_rbar, 1);
        fence_mbar_init();
    }
    float* vec = reinterpret_cast<float*>(base + L.vec);
    float* w1x = vec;
    float* s1 = vec + 3 * a.C1;
    float* t1 = s1 + a.C1;
    float* sl[kMaxTcLayers];
    float* tl[kMaxTcLayers];
    {
        float* p = t1 + 4 * a.C1;
        for (int l = 0; l < a.nl; ++l) { sl[l] = p; tl[l] = p + a.Ntot[l]; p += 2 * a.Ntot[l]; }
    }
    float* w1c = t1 + a.C1;
    // the BN scale of layer 1 is folded into its xyz / centre weights (s (U + d.W) + t = (s U + t) + d.(s W)): one FFMA2 less per channel
    // pair and, for levels without input features, no accumulator to clear
    for (int i = tid; i < 3 * a.C1; i += TcDual::kThreads) {
        const float sc = a.s1 ? __ldg(a.s1 + i % a.C1) : 1.f;
        w1x[i] = __ldg(a.w1x + i) * sc;
        w1c[i] = a.w1c ? __ldg(a.w1c + i) * sc : 0.f;
    }
    for (int i = tid; i < a.C1; i += TcDual::kThreads) { s1[i] = a.s1 ? __ldg(a.s1 + i) : 1.f; t1[i] = __ldg(a.t1 + i); }
    if (tid == 0) { s_nonneg = 1; s_token = 0; }
    __syncthreads();
    for (int l = 0; l < a.nl; ++l)
        for (int i = tid; i < a.Ntot[l]; i += TcDual::kThreads) {
            sl[l][i] = a.s[l] ? __ldg(a.s[l] + i) : 1.f;
            tl[l][i] = __ldg(a.t[l] + i);
            if (l == last && !(sl[l][i] >= 0.f)) s_nonneg = 0;
        }
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int l = 0; l < a.nl; ++l)
            if (!(a.stream_last && l == last)) total += (uint32_t)tc_image_bytes(a.Kd[l], a.Ntot[l], NP);
        if (total) {
            mbar_expect_tx(&s_rbar, total);
            for (int l = 0; l < a.nl; ++l) {
                if (a.stream_last && l == last) continue;
                const uint32_t bytes = (uint32_t)tc_image_bytes(a.Kd[l], a.Ntot[l], NP);
                for (uint32_t o = 0; o < bytes; o += 32768u) bulk_g2s(base + L.w[l] + o, a.image[l] + o, min(32768u, bytes - o), &s_rbar);
            }
            mbar_wait(&s_rbar, 0);
        }
    }
    // ring protocol: one fill outstanding or landed before every issue of the streamed layer
    uint32_t wphase = 0;
    if (issuer && a.stream_last && lane == 0) {
        mbar_expect_tx(&s_wbar[g], L.ring_bytes);
        for (uint32_t o = 0; o < L.ring_bytes; o += 32768u) bulk_g2s(base + L.ring[g] + o, a.image[last] + o, min(32768u, L.ring_bytes - o), &s_wbar[g]);
    }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = warp_uniform(s_tmem) + (uint32_t)g * TcDual::kGroupCols;
    const uint32_t row_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t phase = 0, phase2 = 0;
    constexpr uint32_t a1_col = (DB == 1 || DB == 3) ? 128u : TcDual::A1, a_ps = (DB == 1 || DB == 3) ? 32u : 64u;
    uint8_t* hbuf = base + L.hbuf[g];           // DB == 3: B operand of the transposed last layer
    const bool pool_first = s_nonneg != 0;    // relu(s*d + t) is non-decreasing in d when s >= 0: max over rows commutes with it
    bool have_geo = false;                    // s_geo[g] holds this tile's geometry (written during the previous tile)

    const int G = 128 / a.K;
    const long long ntiles = (a.groups + G - 1) / G;
    if (gtid == 0) s_tile[g][0] = atomicAdd(a.tile_counter, 1u);
    group_bar(g);
    int tpar = 0;
#ifdef PSA_TC_TIMING
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#endif
    for (long long tile = warp_uniform(s_tile[g][0]); tile < ntiles; tile = warp_uniform(s_tile[g][tpar])) {
#ifdef PSA_TC_TIMING
        if (tid == 0) tacc[7] += 1;
#endif
        long long next_tile = 0;
        if (gtid == 0) { const unsigned int t = atomicAdd(a.tile_counter, 1u); s_tile[g][tpar ^ 1] = t; next_tile = t; }
        tpar ^= 1;
        const long long g0 = tile * G;
        const long long gid = g0 + row / a.K;
        const bool valid = gid < a.groups;
        // ---- layer 1 on the FMA pipe, straight into the A operand ----
        {
            float dx = 0.f, dy = 0.f, dz = 0.f;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            const float* urow = nullptr;
            if (a.w1c != nullptr && valid) {
                const float* c = a.new_xyz + (size_t)gid * 3;
                cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
            }
            if (have_geo) {
                const float4 gq = s_geo[g][row];
                dx = gq.x; dy = gq.y; dz = gq.z;
                if (a.uf && valid) urow = a.uf + (size_t)__float_as_int(gq.w) * a.C1;
            } else if (valid) {
                const long long bi = gid / a.m;
                const int j = __ldg(a.idx + gid * a.K + (row % a.K));
                const float* p = a.xyz + ((size_t)bi * a.n + j) * 3;
                const float* c = a.new_xyz + (size_t)gid * 3;
                dx = __ldg(p) - __ldg(c); dy = __ldg(p + 1) - __ldg(c + 1); dz = __ldg(p + 2) - __ldg(c + 2);
                if (a.uf) urow = a.uf + ((size_t)bi * a.n + j) * a.C1;
            }
            const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
            for (int ch = cs; ch < a.C1 / 32; ch += 2) {
                float2 h[16];                         // 32 channels, two per register pair: the chain below runs on the FFMA2 pipe
                const float4* s4 = reinterpret_cast<const float4*>(s1 + ch * 32);
                const float4* t4 = reinterpret_cast<const float4*>(t1 + ch * 32);
                if (urow) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(urow + ch * 32) + q);
                        const float4 sc = s4[q], sh = t4[q];
                        h[2 * q] = __ffma2_rn(make_float2(u.x, u.y), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                        h[2 * q + 1] = __ffma2_rn(make_float2(u.z, u.w), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 sh = t4[q];
                        h[2 * q] = make_float2(sh.x, sh.y); h[2 * q + 1] = make_float2(sh.z, sh.w);
                    }
                }
                if (a.w1c != nullptr) {       // EdgeConv: the part of the first layer that acts on the centre x_i
                    const float2 cx2 = make_float2(cx, cx), cy2 = make_float2(cy, cy), cz2 = make_float2(cz, cz);
                    const float4* cx4 = reinterpret_cast<const float4*>(w1c + ch * 32);
                    const float4* cy4 = reinterpret_cast<const float4*>(w1c + a.C1 + ch * 32);
                    const float4* cz4 = reinterpret_cast<const float4*>(w1c + 2 * a.C1 + ch * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 wx = cx4[q], wy = cy4[q], wz = cz4[q];
                        h[2 * q] = __ffma2_rn(cz2, make_float2(wz.x, wz.y), __ffma2_rn(cy2, make_float2(wy.x, wy.y), __ffma2_rn(cx2, make_float2(wx.x, wx.y), h[2 * q])));
                        h[2 * q + 1] = __ffma2_rn(cz2, make_float2(wz.z, wz.w),
                                                  __ffma2_rn(cy2, make_float2(wy.z, wy.w), __ffma2_rn(cx2, make_float2(wx.z, wx.w), h[2 * q + 1])));
                    }
                }
                const float4* wx4 = reinterpret_cast<const float4*>(w1x + ch * 32);
                const float4* wy4 = reinterpret_cast<const float4*>(w1x + a.C1 + ch * 32);
                const float4* wz4 = reinterpret_cast<const float4*>(w1x + 2 * a.C1 + ch * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 wx = wx4[q], wy = wy4[q], wz = wz4[q];
                    float2 v0 = __ffma2_rn(dz2, make_float2(wz.x, wz.y), __ffma2_rn(dy2, make_float2(wy.x, wy.y), __ffma2_rn(dx2, make_float2(wx.x, wx.y), h[2 * q])));
                    float2 v1 = __ffma2_rn(dz2, make_float2(wz.z, wz.w),
                                           __ffma2_rn(dy2, make_float2(wy.z, wy.w), __ffma2_rn(dx2, make_float2(wx.z, wx.w), h[2 * q + 1])));
                    if (a.relu1) { v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f); }
                    h[2 * q] = v0; h[2 * q + 1] = v1;                    // rows past the end: see affine_chunk
                }
                if constexpr (DB == 3) {
                    if (a.nl == 1) store_h_smem<NP>(hbuf, row, ch * 32, h, ovf);      // layer 1 feeds the transposed layer directly
                    else store_a2_at<NP>(row_taddr + a1_col + ch * 16, h, a_ps, ovf);
                } else {
                    store_a2_at<NP>(row_taddr + a1_col + ch * 16, h, a_ps, ovf);
                }
            }
        }
        if constexpr (DB == 3) fence_proxy_async_smem();
        tmem_st_wait();
        fence_before_thread_sync();
        group_bar(g);
        TC_STAMP(0);
        for (int l = 0; l < a.nl; ++l) {
            const int NT = (int)warp_uniform((uint32_t)a.Ntot[l]) / kNt, KC = (int)warp_uniform((uint32_t)a.Kd[l]) / 64;
            if (l != last) {
                // inner layer, one or two 64-wide tiles: the next A operand may only be written once ALL MMAs of this
                // layer are done reading the current one, so the first tile's activations wait in registers
                float2 h0[16], h1[16];
                for (int nt = 0; nt < NT; ++nt) {
                    if (issuer) {
                        pipe_acquire(&s_token, lane);
                        fence_after_thread_sync();
                        issue_tile<NP, DB>(tmem_base, smem_u32(base + L.w[l]) + (uint32_t)nt * KC * tc_block_bytes(kNt, NP), KC, 0);
                        mma_commit(&s_mbar[g]);
                        pipe_release(&s_token, lane);
                    }
                    TC_STAMP(1);
                    mbar_wait(&s_mbar[g], phase);
                    phase ^= 1u;
                    fence_after_thread_sync();
                    TC_STAMP(2);
                    uint32_t d[32];
                    tmem_ld32(row_taddr + TcDual::D + cs * 32, d);
                    tmem_ld_wait();
                    if (nt == 0) affine_chunk(d, sl[l] + cs * 32, tl[l] + cs * 32, a.relu[l], h0);
                    else affine_chunk(d, sl[l] + kNt + cs * 32, tl[l] + kNt + cs * 32, a.relu[l], h1);
                    if (nt + 1 < NT) { fence_before_thread_sync(); group_bar(g); }       // D drained before the next tile lands in it
                }
                if constexpr (DB == 3) {
                    store_h_smem<NP>(hbuf, row, cs * 32, h0, ovf);                    // NT == 1: the inner layer of this mode is 64 wide
                    fence_proxy_async_smem();
                } else {
                    store_a2_at<NP>(row_taddr + a1_col + cs * 16, h0, a_ps, ovf);
                    if (NT == 2) store_a2_at<NP>(row_taddr + a1_col + (2 + cs) * 16, h1, a_ps, ovf);
                }
                tmem_st_wait();
                fence_before_thread_sync();
                group_bar(g);
                TC_STAMP(3);
            } else if constexpr (DB == 3) {
                // ---- last layer, transposed: D^T[128 ch][128 rows] = W^T . H^T, both operands from shared memory ----
                if (issuer) {
                    pipe_acquire(&s_token, lane);
                    fence_after_thread_sync();
                    const uint32_t idesc = make_idesc(Split<NP>::kFmt, 128, 128);
                    const SmemDescBase wa = smem_desc_base(warp_uniform(smem_u32(base + L.w[l])));
                    const SmemDescBase hb = smem_desc_base(warp_uniform(smem_u32(hbuf)));
                    const uint32_t dT = warp_uniform(tmem_base);
#pragma unroll
                    for (int t = 0; t < Split<NP>::kTerms; ++t)
#pragma unroll
                        for (int s4 = 0; s4 < 4; ++s4)
                            mma_bf16_ss(dT, smem_desc_at(wa, Split<NP>::w(t) * 16384u + s4 * 32), smem_desc_at(hb, Split<NP>::a(t) * 16384u + s4 * 32), idesc,
                                        (t | s4) ? 1u : 0u);
                    mma_commit(&s_mbar[g]);
                    pipe_release(&s_token, lane);
                }
                TC_STAMP(4);
                {
                    // the next tile's index -> point -> offset chain (two dependent L2 round trips) runs under these MMAs
                    const long long ntile = (long long)s_tile[g][tpar];
                    have_geo = ntile < ntiles;
                    if (have_geo && cs == 0) {
                        const long long ngid = ntile * G + row / a.K;
                        float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ngid < a.groups) {
                            const long long bi = ngid / a.m;
                            const int j = __ldg(a.idx + ngid * a.K + (row % a.K));
                            const float* p = a.xyz + ((size_t)bi * a.n + j) * 3;
                            const float* c = a.new_xyz + (size_t)ngid * 3;
                            gq.x = __ldg(p) - __ldg(c); gq.y = __ldg(p + 1) - __ldg(c + 1); gq.z = __ldg(p + 2) - __ldg(c + 2);
                            gq.w = __int_as_float((int)(bi * a.n + j));
                        }
                        s_geo[g][row] = gq;
                    }
                }
                mbar_wait(&s_mbar[g], phase);
                phase ^= 1u;
                fence_after_thread_sync();
                TC_STAMP(5);
                {
                    // lane = channel (32 quarter + lane); this warp's half of the rows = columns [64 cs, 64 cs + 64): 64 / K neighbourhoods
                    const int chn = quarter * 32 + lane;
                    const float sc = sl[l][chn], sh = tl[l][chn];
                    const int npg = 64 / a.K;                                        // K = 32: two neighbourhoods per half, K = 64: one
                    for (int j = 0; j < npg; ++j) {
                        float mx = -FLT_MAX;
                        for (int c = 0; c < a.K / 32; ++c) {
                            uint32_t d[32];
                            tmem_ld32(row_taddr + (uint32_t)(cs * 64 + j * a.K + c * 32), d);
                            tmem_ld_wait();
                            if (pool_first) {
#pragma unroll
                                for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(d[q]));
                            } else {
#pragma unroll
                                for (int q = 0; q < 32; ++q) {
                                    float v = fmaf(__uint_as_float(d[q]), sc, sh);
                                    if (a.relu[l]) v = fmaxf(v, 0.f);
                                    mx = fmaxf(mx, v);
                                }
                            }
                        }
                        if (pool_first) {
                            mx = fmaf(mx, sc, sh);
                            if (a.relu[l]) mx = fmaxf(mx, 0.f);
                        }
                        const long long og = g0 + cs * npg + j;
                        if (og < a.groups) a.out[(size_t)og * a.Ntot[l] + chn] = mx;      // 32 lanes = 128 contiguous bytes
                    }
                }
                fence_before_thread_sync();
                group_bar(g);
                TC_STAMP(6);
            } else {
                const bool streamed = a.stream_last != 0;
                const int quarters_per_group = a.K / 32;        // 1, 2 or 4
                const long long wg = g0 + (quarter * 32) / a.K; // this warp's neighbourhood
                // pooled epilogue of output tile `nt` out of D slot `dslot`
                auto pooled_epilogue = [&](int nt, int dslot) {
                    uint32_t d[32];
                    tmem_ld32(row_taddr + (dslot == 0 ? dual_dcol<DB>(0) : dual_dcol<DB>(1)) + cs * 32, d);
                    tmem_ld_wait();
                    float v[32];
                    if (pool_first) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(d[q]);
                    } else {
                        float2 v2[16];
                        affine_chunk(d, sl[l] + nt * kNt + cs * 32, tl[l] + nt * kNt + cs * 32, a.relu[l], v2);
#pragma unroll
                        for (int q = 0; q < 16; ++q) { v[2 * q] = v2[q].x; v[2 * q + 1] = v2[q].y; }
                    }
                    float mx = warp_colmax_32x32(v, lane);
                    if (quarters_per_group > 1) {
                        s_red[warp][lane] = mx;
                        group_bar(g);
                        if ((quarter % quarters_per_group) == 0)
                            for (int o = 1; o < quarters_per_group; ++o) mx = fmaxf(mx, s_red[warp + o][lane]);
                    }
                    if (pool_first) {
                        mx = fmaf(mx, sl[l][nt * kNt + cs * 32 + lane], tl[l][nt * kNt + cs * 32 + lane]);
                        if (a.relu[l]) mx = fmaxf(mx, 0.f);
                    }
                    if ((quarter % quarters_per_group) == 0 && wg < a.groups)
                        a.out[(size_t)wg * a.Ntot[l] + nt * kNt + cs * 32 + lane] = mx;
                };
                constexpr int kStep = DBUF ? 2 : 1;
                for (int nt = 0; nt < NT; nt += kStep) {
                    if (issuer) {
                        if (streamed) { mbar_wait(&s_wbar[g], wphase); wphase ^= 1u; }
                        pipe_acquire(&s_token, lane);   // (ends in __syncwarp: the issue below must be warp-uniform)
                        fence_after_thread_sync();
                        const uint32_t blocks = streamed ? smem_u32(base + L.ring[g]) : smem_u32(base + L.w[l]) + (uint32_t)nt * KC * tc_block_bytes(kNt, NP);
                        issue_tile<NP, DB>(tmem_base, blocks, KC, 0);
                        mma_commit(&s_mbar[g]);
                        if (DBUF) {                 // the pair's second tile goes to the other D slot right away (NT is even)
                            issue_tile<NP, DB>(tmem_base, blocks + (uint32_t)KC * tc_block_bytes(kNt, NP), KC, 1);
                            mma_commit(&s_mbar2[g]);
                        }
                        pipe_release(&s_token, lane);
                    }
                    TC_STAMP(4);
                    if (nt + kStep >= NT) {
                        // the next tile's index -> point -> offset chain (two dependent L2 round trips) runs under these MMAs
                        const long long ntile = (long long)s_tile[g][tpar];
                        have_geo = ntile < ntiles;
                        if (have_geo && cs == 0) {
                            const long long ngid = ntile * G + row / a.K;
                            float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ngid < a.groups) {
                                const long long bi = ngid / a.m;
                                const int j = __ldg(a.idx + ngid * a.K + (row % a.K));
                                const float* p = a.xyz + ((size_t)bi * a.n + j) * 3;
                                const float* c = a.new_xyz + (size_t)ngid * 3;
                                gq.x = __ldg(p) - __ldg(c); gq.y = __ldg(p + 1) - __ldg(c + 1); gq.z = __ldg(p + 2) - __ldg(c + 2);
                                gq.w = __int_as_float((int)(bi * a.n + j));
                            }
                            s_geo[g][row] = gq;       // read after this tile's closing group barriers
                        }
                    }
                    mbar_wait(&s_mbar[g], phase);
                    phase ^= 1u;
                    fence_after_thread_sync();
                    TC_STAMP(5);
                    if (issuer && streamed) {
                        // the ring is free again: fetch the next tile (wrapping to tile 0 for the group's next row tile)
                        const bool more = (nt + 1 < NT) || (__shfl_sync(0xffffffffu, next_tile, 0) < ntiles);
                        if (more && lane == 0) {
                            const int rt = (nt + 1) % NT;
                            mbar_expect_tx(&s_wbar[g], L.ring_bytes);
                            for (uint32_t o = 0; o < L.ring_bytes; o += 32768u)
                                bulk_g2s(base + L.ring[g] + o, a.image[l] + (size_t)rt * L.ring_bytes + o, min(32768u, L.ring_bytes - o), &s_wbar[g]);
                        }
                    }
                    pooled_epilogue(nt, 0);
                    if (DBUF) {
                        if (quarters_per_group > 1) group_bar(g);       // s_red is reused by the second tile
                        mbar_wait(&s_mbar2[g], phase2);
                        phase2 ^= 1u;
                        fence_after_thread_sync();
                        pooled_epilogue(nt + 1, 1);
                    }
                    // D fully read (and s_red consumed) by every warp of the group before the next MMAs / maxima land
                    fence_before_thread_sync();
                    group_bar(g);
                    TC_STAMP(6);
                }
            }
        }
    }
#ifdef PSA_TC_TIMING
    if (tid == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&g_tc_timing[i], tacc[i]);
#endif
    if constexpr (NP == 2) {
        if (f16x2_overflowed(ovf)) atomicOr(a.ovf, 1u);
        if (tid == 0)
            for (int l = 0; l < a.nl; ++l)
                if (a.wflag[l] != nullptr && *a.wflag[l] != 0u) atomicOr(a.ovf, 1u);
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Dense layer on the tensor cores: out = relu?((x . W) * scale + shift), optional max over runs of pool_k rows.
// CTA = 128 rows x Nt output channels.  K is walked in segments of 128: per segment the row warps quantise their x
// chunk into the TMEM A operand, one thread issues the three-term MMAs against the segment's weight blocks (bulk-copied
// into a two-slot shared-memory ring one segment ahead), and the segment's D is added to fp32 register accumulators --
// so the truncating TMEM accumulation never runs over more than 128 K, however long the dot product is.
// ------------------------------------------------------------------------------------------------------------------
struct TcDenseArgs {
    long long rows;
    int K, Kp, N;          // Kp = K rounded up to 64
    int pool_k;            // 1, 32, 64 or 128
    int relu;
    const float* x;        // (rows, K)
    const uint8_t* image;  // weight image for (Kp, N)
    const float* scale;    // (N) or null
    const float* shift;    // (N) or null
    const float* xyz3;     // optional side input (rows, 3): out += xyz3 . w3 before scale/shift (the xyz rows of a
    const float* w3;       // (3, N)                          [xyz, features] . W product, kept off the K loop)
    float* out;
    // training-mode forward (tc_dense3 only): the input is relu(x * in_scale + in_shift) per INPUT channel (the previous layer's
    // batch norm, applied while the operand is staged) and per-tile column statistics of the stored values are written
    const float* in_scale = nullptr;   // (K) or null
    const float* in_shift = nullptr;
    int in_relu = 0;
    float* stat_partial = nullptr;     // (row tiles, 2, N) or null
    // operand split (see TcArgs): np = 2 launches raise *ovf, the np = 3 rerun is skipped unless *run_if != 0
    unsigned int* ovf = nullptr;
    const unsigned int* run_if = nullptr;
    const unsigned int* wflag = nullptr;
};

// ------------------------------------------------------------------------------------------------------------------
// tc_dense2_kernel -- the dense layer as a software pipeline.
//   CTA = 128 rows x Nt output channels (Nt = 128, or 64 for a 64-wide layer), 17 warps, one CTA per SM:
//   * 16 ROW warps (quarter = w & 3 -> TMEM lanes, slot = w >> 2 -> 32-column chunk) quantise their chunk of the next
//     K = 128 segment of x into NP pieces in one of TWO A-operand buffers in TMEM, while
//   * the ISSUER warp (converged, tc_common.cuh) waits for that buffer, the segment's weight blocks (cp.async.bulk
//     into a two-slot shared-memory ring, refilled as soon as the MMAs that read a slot have completed) and for D to
//     be drained, then issues the segment's MMAs (three or six terms) and commits;
//   * the row warps add each segment's D to fp32 register accumulators (the truncating TMEM accumulation never runs
//     over more than 128 K), then run the epilogue (affine, ReLU, xyz side input, max-pool).
//   A-preparation of segment s+1 overlaps the MMAs of segment s; all hand-offs are mbarriers, no CTA-wide barrier
//   inside the K loop.   TMEM: D 128 | A[0] 192 | A[1] 192 columns.
// ------------------------------------------------------------------------------------------------------------------
struct TcDense2 {
    static constexpr int kRowWarps = 16, kThreads = 17 * 32;
    static constexpr uint32_t D = 0, A0 = 128, ABUF = 192;     // A buffer b at A0 + b * ABUF, pieces 64 columns apart
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }

template <int NT_, int NP>
__global__ void __launch_bounds__(TcDense2::kThreads, 1)
tc_dense2_kernel(const __grid_constant__ TcDenseArgs a) {
    if (a.run_if != nullptr && *a.run_if == 0u) return;
    uint32_t ovf = 0u;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_wfull[2];     // weight slot landed (tx)
    __shared__ __align__(8) uint64_t s_afull[2];     // A buffer written by all 16 row warps
    __shared__ __align__(8) uint64_t s_mma;          // segment's MMAs complete
    __shared__ __align__(8) uint64_t s_dfree;        // D drained by all 16 row warps
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[TcDense2::kRowWarps][32];
    __shared__ __align__(16) float s_vec[5][NT_];    // scale, shift, three xyz rows of W for this CTA's output channels
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp_u = (int)warp_uniform((uint32_t)(tid >> 5));
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KCtot = a.Kp / 64;
    constexpr uint32_t bb = NT_ * 128u * NP, slot_bytes = 2u * bb;
    const int nt = blockIdx.y;
    for (int i = tid; i < NT_; i += TcDense2::kThreads) {
        const int c = nt * NT_ + i;
        s_vec[0][i] = a.scale ? __ldg(a.scale + c) : 1.f;
        s_vec[1][i] = a.shift ? __ldg(a.shift + c) : 0.f;
        s_vec[2][i] = a.xyz3 ? __ldg(a.w3 + c) : 0.f;
        s_vec[3][i] = a.xyz3 ? __ldg(a.w3 + a.N + c) : 0.f;
        s_vec[4][i] = a.xyz3 ? __ldg(a.w3 + 2 * a.N + c) : 0.f;
    }
    const long long row0 = (long long)blockIdx.x * 128;
    const int nseg = (KCtot + 1) / 2;
    const uint8_t* img = a.image + (size_t)nt * KCtot * bb;

    if (warp_u == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        mbar_init(&s_wfull[0], 1); mbar_init(&s_wfull[1], 1);
        mbar_init(&s_afull[0], TcDense2::kRowWarps); mbar_init(&s_afull[1], TcDense2::kRowWarps);
        mbar_init(&s_mma, 1); mbar_init(&s_dfree, TcDense2::kRowWarps);
        fence_mbar_init();
    }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = warp_uniform(s_tmem);
#ifdef PSA_TC_TIMING
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 1};
    long long tprev = clock64();
#endif

    if (warp_u == TcDense2::kRowWarps) {
        // ================= issuer warp =================
        auto load_seg = [&](int sg) {       // one lane
            const int kcs = min(2, KCtot - 2 * sg);
            const uint32_t bytes = (uint32_t)kcs * bb;
            uint64_t* bar = &s_wfull[sg & 1];
            mbar_expect_tx(bar, bytes);
            for (uint32_t o = 0; o < bytes; o += 32768u)
                bulk_g2s(base + (sg & 1) * slot_bytes + o, img + (size_t)sg * slot_bytes + o, min(32768u, bytes - o), bar);
        };
        if (lane == 0) { load_seg(0); if (nseg > 1) load_seg(1); }
        __syncwarp();
        for (int sg = 0; sg < nseg; ++sg) {
            const int b = sg & 1;
            const uint32_t par = (uint32_t)((sg >> 1) & 1);
            mbar_wait(&s_wfull[b], par);
            mbar_wait(&s_afull[b], par);
            if (sg > 0) {
                mbar_wait(&s_dfree, (uint32_t)((sg - 1) & 1));        // D drained => MMAs of segment sg-1 completed too
                if (lane == 0 && sg + 1 < nseg) load_seg(sg + 1);     // their weight slot is free again
            }
            __syncwarp();
            fence_after_thread_sync();
            const uint32_t blocks = smem_u32(base) + (uint32_t)b * slot_bytes;
            const uint32_t a1 = TcDense2::A0 + (uint32_t)b * TcDense2::ABUF;
            if (KCtot - 2 * sg >= 2) issue_tile_c<2, NT_, 64, NP>(tmem_base, TcDense2::D, a1, blocks);
            else issue_tile_c<1, NT_, 64, NP>(tmem_base, TcDense2::D, a1, blocks);
            mma_commit(&s_mma);
        }
    } else {
        // ================= row warps =================
        const int quarter = warp_u & 3, cs = warp_u >> 2;
        const int row = quarter * 32 + lane;
        const long long grow = row0 + row;
        const bool valid = grow < a.rows;
        const uint32_t row_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const bool vec_ok = ((a.K & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
        const bool has_out_chunk = cs < NT_ / 32;
        float acc[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) acc[q] = 0.f;

        // this warp's 32-wide chunk of segment sg of x, raw (loads only: issued early, consumed a segment later)
        auto load_x = [&](int sg, float (&h)[32]) {
            const int kcs = min(2, KCtot - 2 * sg);
            if (cs < kcs * 2) {
                const int k0 = sg * 128 + cs * 32;
                const float* xr = a.x + (size_t)(valid ? grow : 0) * a.K + k0;
                if (valid && vec_ok && k0 + 32 <= a.K) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(xr) + q);
                        h[4 * q] = u.x; h[4 * q + 1] = u.y; h[4 * q + 2] = u.z; h[4 * q + 3] = u.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) h[q] = (valid && k0 + q < a.K) ? __ldg(xr + q) : 0.f;
                }
            }
        };
        // split into NP pieces -> A buffer sg & 1, then tell the issuer
        auto store_x = [&](int sg, float (&h)[32]) {
            const int kcs = min(2, KCtot - 2 * sg);
            if (cs < kcs * 2) {
                store_a_at<NP>(row_taddr + TcDense2::A0 + (uint32_t)(sg & 1) * TcDense2::ABUF + cs * 16, h, 64, ovf);
                tmem_st_wait();
            }
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_afull[sg & 1]);
        };

        TC_STAMP(0);
        float h[32];
        load_x(0, h);
        store_x(0, h);
        if (nseg > 1) load_x(1, h);
        TC_STAMP(1);
        for (int sg = 0; sg < nseg; ++sg) {
            if (sg + 1 < nseg) {
                store_x(sg + 1, h);                   // buffer (sg+1)&1 was released by the MMAs of segment sg-1 (waited below)
                if (sg + 2 < nseg) load_x(sg + 2, h); // in flight across this segment's MMA wait and drain
            }
            TC_STAMP(2);
            mbar_wait(&s_mma, (uint32_t)(sg & 1));
            fence_after_thread_sync();
            TC_STAMP(3);
            if (has_out_chunk) {
                uint32_t d[32];
                tmem_ld32(row_taddr + TcDense2::D + cs * 32, d);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; ++q) acc[q] += __uint_as_float(d[q]);
            }
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_dfree);
            TC_STAMP(4);
        }
        // ---- epilogue ----
        const int col0 = nt * NT_ + cs * 32;
        float v[32];
        if (has_out_chunk) {
            float sx3 = 0.f, sy3 = 0.f, sz3 = 0.f;
            if (a.xyz3 != nullptr && valid) {
                sx3 = __ldg(a.xyz3 + (size_t)grow * 3); sy3 = __ldg(a.xyz3 + (size_t)grow * 3 + 1); sz3 = __ldg(a.xyz3 + (size_t)grow * 3 + 2);
            }
            const float* vs = &s_vec[0][cs * 32];
            if (a.xyz3 != nullptr) {
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    acc[q] = fmaf(sz3, vs[4 * NT_ + q], fmaf(sy3, vs[3 * NT_ + q], fmaf(sx3, vs[2 * NT_ + q], acc[q])));
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                float x = fmaf(acc[q], vs[q], vs[NT_ + q]);
                if (a.relu) x = fmaxf(x, 0.f);
                v[q] = x;
            }
        }
        if (a.pool_k == 1) {
            if (has_out_chunk && valid) {
                float4* o = reinterpret_cast<float4*>(a.out + (size_t)grow * a.N + col0);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        } else {
            const bool big = a.pool_k > 128;                         // the whole 128-row tile lies inside one group
            const int quarters_per_group = big ? 4 : a.pool_k / 32;
            const long long wg = (row0 + quarter * 32) / a.pool_k;
            float mx = -FLT_MAX;
            if (has_out_chunk) {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = valid ? v[q] : -FLT_MAX;
                mx = warp_colmax_32x32(v, lane);
            }
            if (quarters_per_group > 1) {
                s_red[warp_u][lane] = mx;
                named_bar_sync(1, TcDense2::kRowWarps * 32);
                if ((quarter % quarters_per_group) == 0)
                    for (int o = 1; o < quarters_per_group; ++o) mx = fmaxf(mx, s_red[warp_u + o][lane]);
            }
            if (has_out_chunk && (quarter % quarters_per_group) == 0 && row0 + quarter * 32 < a.rows) {
                if (!big) {
                    a.out[(size_t)wg * a.N + col0 + lane] = mx;
                } else {
                    int code = __float_as_int(mx);
                    code = code >= 0 ? code : code ^ 0x7fffffff;
                    atomicMax(reinterpret_cast<int*>(a.out) + (size_t)wg * a.N + col0 + lane, code);
                }
            }
        }
    }
    TC_STAMP(5);
    if constexpr (NP == 2) {
        if (f16x2_overflowed(ovf) || (tid == 0 && a.wflag != nullptr && *a.wflag != 0u)) atomicOr(a.ovf, 1u);
    }
    fence_before_thread_sync();
    __syncthreads();
    TC_STAMP(6);
#ifdef PSA_TC_TIMING
    if (tid == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&g_tc_timing[i], tacc[i]);
#endif
    if (warp_u == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// tc_dense3_kernel -- the dense layer in TRANSPOSED form, both operands from shared memory:
//      D^T[channel][row] = sum_k W^T[channel][k] . x^T[k][row]
//   M = 128 output channels (TMEM lanes), N = 128 rows (TMEM columns), K walked in 64-wide blocks, two stages.
//   * A operand = the weight image block exactly as tc_dense2 uses it as B ([128 ch][64 k] K-major SWIZZLE_128B, three
//     bf16 pieces, 48 KB, one cp.async.bulk);
//   * B operand = x quantised into the same layout by 16 prep warps: each warp reads 8 rows x 256 B fully COALESCED
//     (lane = two consecutive k), splits into NP pieces and writes 4-byte words into the swizzled rows -- conflict-free.
//     (tc_dense2's lane = row loads cost 4096 L1 wavefronts per K = 128 segment; here 256 per 64-K block.)
//   * the issuer warp waits for a stage's two operands and issues its 24 MMAs; the commit frees the stage;
//   * accumulation: K-block kb goes to accumulator kb & 3 (4 x 128 TMEM columns), so no accumulator takes more than 48
//     truncating adds at K = 512 (the bound tc_dense2 keeps with its register sums); the epilogue adds the four;
//   * epilogue with lane = CHANNEL: scale / shift / xyz-side weights are per-thread scalars, a max over rows is an
//     in-thread reduction (no shuffles), and every global store is a coalesced 128-byte row segment.
// ------------------------------------------------------------------------------------------------------------------
struct TcDense3 {
    static constexpr int kPrepWarps = 16, kThreads = 17 * 32;
    static constexpr uint32_t kPiece = 128u * 128u;          // one 16-bit piece of a 128 x 64 block: 16 KB
};

// CP = channel tiles per CTA.  CP = 2 (fp16x2, N a multiple of 256): the x block of a K step is staged ONCE and multiplied with the
// weight blocks of two 128-channel tiles (accumulators 0/1 and 2/3) -- the staging work per output halves and SA3's 512 -> 1024 layer
// becomes a single wave of 128 CTAs instead of 1.7 waves of 256.
template <int NP, int CP>
__global__ void __launch_bounds__(TcDense3::kThreads, 1)
tc_dense3_kernel(const __grid_constant__ TcDenseArgs a) {
    if (a.run_if != nullptr && *a.run_if == 0u) return;
    constexpr uint32_t kBlock = NP * TcDense3::kPiece;       // 48 KB (bf16x3) or 32 KB (fp16x2)
    constexpr int kAcc = 4 / CP;                             // TMEM accumulators (128 columns each) per channel tile
    uint32_t ovf = 0u;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_wfull[2];     // weight block landed (tx)
    __shared__ __align__(8) uint64_t s_xfull[2];     // x block written by the 16 prep warps
    __shared__ __align__(8) uint64_t s_free[2];      // the MMAs that read a stage completed
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[TcDense3::kPrepWarps][32];
    __shared__ float s_xyz[128][3];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp_u = (int)warp_uniform((uint32_t)(tid >> 5));
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* wslot = base;                                   // 2 stages x CP weight blocks
    uint8_t* xslot = base + 2 * CP * kBlock;                 // 2 stages x 1 x block
    const int KB = a.Kp / 64;
    const int nt = blockIdx.y * CP;                          // first channel tile of this CTA
    const long long row0 = (long long)blockIdx.x * 128;
    const uint8_t* img = a.image + (size_t)nt * KB * kBlock;

    if (warp_u == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        mbar_init(&s_wfull[0], 1); mbar_init(&s_wfull[1], 1);
        mbar_init(&s_xfull[0], TcDense3::kPrepWarps); mbar_init(&s_xfull[1], TcDense3::kPrepWarps);
        mbar_init(&s_free[0], 1); mbar_init(&s_free[1], 1);
        fence_mbar_init();
    }
    if (a.xyz3 != nullptr)
        for (int i = tid; i < 128 * 3; i += TcDense3::kThreads) {
            const long long r = row0 + i / 3;
            s_xyz[i / 3][i % 3] = r < a.rows ? __ldg(a.xyz3 + (size_t)r * 3 + i % 3) : 0.f;
        }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = warp_uniform(s_tmem);
#ifdef PSA_TC_TIMING
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 1};
    long long tprev = clock64();
#endif

    if (warp_u == TcDense3::kPrepWarps) {
        // ================= issuer warp =================
        const uint32_t idesc = make_idesc(Split<NP>::kFmt, 128, 128);
        for (int kb = 0; kb < KB; ++kb) {
            const int st = kb & 1;
            const uint32_t par = (uint32_t)((kb >> 1) & 1);
            mbar_wait(&s_wfull[st], par);
            mbar_wait(&s_xfull[st], par);
            __syncwarp();
            fence_after_thread_sync();
            const SmemDescBase xb = smem_desc_base(warp_uniform(smem_u32(xslot) + (uint32_t)st * kBlock));
            const uint32_t first = kb < kAcc ? 0u : 1u;      // an accumulator's first block overwrites it
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const uint32_t d = tmem_base + (uint32_t)(c * kAcc + (kb % kAcc)) * 128u;
                const SmemDescBase wa = smem_desc_base(warp_uniform(smem_u32(wslot) + (uint32_t)(st * CP + c) * kBlock));
#pragma unroll
                for (int t = 0; t < Split<NP>::kTerms; ++t)
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4)
                        mma_bf16_ss(d, smem_desc_at(wa, Split<NP>::w(t) * TcDense3::kPiece + s4 * 32),
                                    smem_desc_at(xb, Split<NP>::a(t) * TcDense3::kPiece + s4 * 32), idesc, (t | s4) ? 1u : first);
            }
            mma_commit(&s_free[st]);
        }
    } else {
        // ================= prep warps: x block -> B operand; then the epilogue =================
        const bool vec2 = ((a.K & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 7) == 0);
        float2 v[8];
        // this warp's 8 rows x 64 k of block kb, raw: issued a block ahead, before the stage is known to be free
        auto load_x = [&](int kb) {
            const int k = kb * 64 + 2 * lane;
            float2 isc = make_float2(1.f, 1.f), ish = make_float2(0.f, 0.f);
            if (a.in_scale != nullptr) {
                if (k < a.K) { isc.x = __ldg(a.in_scale + k); ish.x = __ldg(a.in_shift + k); }
                if (k + 1 < a.K) { isc.y = __ldg(a.in_scale + k + 1); ish.y = __ldg(a.in_shift + k + 1); }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long r = row0 + warp_u * 8 + i;
                v[i] = make_float2(0.f, 0.f);
                if (r < a.rows) {
                    const float* xr = a.x + (size_t)r * a.K + k;
                    if (vec2 && k + 1 < a.K) v[i] = __ldg(reinterpret_cast<const float2*>(xr));
                    else { if (k < a.K) v[i].x = __ldg(xr); if (k + 1 < a.K) v[i].y = __ldg(xr + 1); }
                    if (a.in_scale != nullptr) {       // previous layer's batch norm (+ relu) on the fly; padded rows / channels stay 0
                        if (k < a.K) { v[i].x = fmaf(v[i].x, isc.x, ish.x); if (a.in_relu) v[i].x = fmaxf(v[i].x, 0.f); }
                        if (k + 1 < a.K) { v[i].y = fmaf(v[i].y, isc.y, ish.y); if (a.in_relu) v[i].y = fmaxf(v[i].y, 0.f); }
                    }
                }
            }
        };
        load_x(0);
        TC_STAMP(0);
        for (int kb = 0; kb < KB; ++kb) {
            const int st = kb & 1;
            if (kb >= 2) mbar_wait(&s_free[st], (uint32_t)(((kb - 2) >> 1) & 1));      // stage released by block kb-2's MMAs
            TC_STAMP(1);
            if (warp_u == 0 && lane == 0) {                                             // its weight slot is free too
                mbar_expect_tx(&s_wfull[st], CP * kBlock);
                for (int c = 0; c < CP; ++c)
                    for (uint32_t o = 0; o < kBlock; o += 16384u)
                        bulk_g2s(wslot + (uint32_t)(st * CP + c) * kBlock + o, img + ((size_t)c * KB + kb) * kBlock + o, 16384u, &s_wfull[st]);
            }
            uint8_t* xs = xslot + (uint32_t)st * kBlock;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t rr = (uint32_t)(warp_u * 8 + i);
                const uint32_t off = swz_off_bf16(rr, 2u * (uint32_t)lane, 128u);
                uint32_t pc[NP];
                split_pair<NP>(v[i].x, v[i].y, pc, ovf);
#pragma unroll
                for (int j = 0; j < NP; ++j) *reinterpret_cast<uint32_t*>(xs + (uint32_t)j * TcDense3::kPiece + off) = pc[j];
            }
            fence_proxy_async_smem();            // generic-proxy writes -> visible to the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_xfull[st]);
            if (kb + 1 < KB) load_x(kb + 1);
            TC_STAMP(2);
        }
        // all MMAs done: the last block's commit covers every earlier one (in-order completion)
        mbar_wait(&s_free[(KB - 1) & 1], (uint32_t)(((KB - 1) >> 1) & 1));
        fence_after_thread_sync();
        TC_STAMP(3);

        // ---- epilogue: lane = channel ----
        const int quarter = warp_u & 3, slot = warp_u >> 2;          // channels 32*quarter.., rows 32*slot..
        const long long rbase = row0 + slot * 32;
#pragma unroll 1
        for (int ct = 0; ct < CP; ++ct) {                             // the CTA's channel tiles, one after the other
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ct * kAcc) * 128u + (uint32_t)slot * 32u;
        const int ch = (nt + ct) * 128 + quarter * 32 + lane;
        float acc[32];
        {
            uint32_t d[32];
            tmem_ld32(taddr, d);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(d[q]);
            const int nacc = KB < kAcc ? KB : kAcc;
            for (int c = 1; c < nacc; ++c) {
                tmem_ld32(taddr + (uint32_t)c * 128u, d);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; ++q) acc[q] += __uint_as_float(d[q]);
            }
        }
        const float sc = a.scale ? __ldg(a.scale + ch) : 1.f;
        const float sh = a.shift ? __ldg(a.shift + ch) : 0.f;
        if (a.xyz3 != nullptr) {
            const float w0 = __ldg(a.w3 + ch), w1 = __ldg(a.w3 + a.N + ch), w2 = __ldg(a.w3 + 2 * a.N + ch);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float* xq = s_xyz[slot * 32 + q];
                acc[q] = fmaf(xq[2], w2, fmaf(xq[1], w1, fmaf(xq[0], w0, acc[q])));
            }
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            float x = fmaf(acc[q], sc, sh);
            if (a.relu) x = fmaxf(x, 0.f);
            acc[q] = x;
        }
        if (a.pool_k == 1) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if (rbase + q < a.rows) a.out[(size_t)(rbase + q) * a.N + ch] = acc[q];       // 32 lanes = 128 contiguous bytes
            if (a.stat_partial != nullptr) {
                // column statistics of the stored tile: lane = channel, so sum / sum of squares over this slot's 32 rows are
                // in-thread; the four row slots of a channel quarter fold through shared memory in slot order (deterministic)
                float ssum = 0.f, ssq = 0.f;
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (rbase + q < a.rows) { ssum += acc[q]; ssq = fmaf(acc[q], acc[q], ssq); }
                __shared__ float s_st[2][TcDense3::kPrepWarps][32];
                s_st[0][warp_u][lane] = ssum;
                s_st[1][warp_u][lane] = ssq;
                named_bar_sync(1, TcDense3::kPrepWarps * 32);
                if (slot == 0) {
                    float t0 = 0.f, t1 = 0.f;
#pragma unroll
                    for (int o = 0; o < 4; ++o) { t0 += s_st[0][quarter + 4 * o][lane]; t1 += s_st[1][quarter + 4 * o][lane]; }
                    float* dst = a.stat_partial + (size_t)blockIdx.x * 2 * a.N;
                    dst[ch] = t0;
                    dst[a.N + ch] = t1;
                }
            }
        } else {
            float mx = -FLT_MAX;
#pragma unroll
            for (int q = 0; q < 32; ++q) mx = (rbase + q < a.rows) ? fmaxf(mx, acc[q]) : mx;
            const bool big = a.pool_k > 128;
            const int slots_per_group = big ? 4 : a.pool_k / 32;         // 1, 2 or 4 row slots per pooling group
            if (slots_per_group > 1) {
                s_red[warp_u][lane] = mx;
                named_bar_sync(1, TcDense3::kPrepWarps * 32);
                if ((slot % slots_per_group) == 0)
                    for (int o = 1; o < slots_per_group; ++o) mx = fmaxf(mx, s_red[warp_u + 4 * o][lane]);
            }
            if ((slot % slots_per_group) == 0 && rbase < a.rows) {
                const long long wg = rbase / a.pool_k;
                if (!big) {
                    a.out[(size_t)wg * a.N + ch] = mx;
                } else {
                    int code = __float_as_int(mx);
                    code = code >= 0 ? code : code ^ 0x7fffffff;
                    atomicMax(reinterpret_cast<int*>(a.out) + (size_t)wg * a.N + ch, code);
                }
            }
        }
        if (CP > 1) named_bar_sync(1, TcDense3::kPrepWarps * 32);       // s_red / s_st are reused by the next channel tile
        }
    }
    TC_STAMP(4);
    if constexpr (NP == 2) {
        if (f16x2_overflowed(ovf) || (tid == 0 && a.wflag != nullptr && *a.wflag != 0u)) atomicOr(a.ovf, 1u);
    }
    fence_before_thread_sync();
    __syncthreads();
    TC_STAMP(5);
#ifdef PSA_TC_TIMING
    if (tid == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&g_tc_timing[i], tacc[i]);
#endif
    if (warp_u == 0) tmem_dealloc(tmem_base, 512);
}

bool tc_dense_eligible(long long rows, int K, int N, int pool_k) {
    if (rows < 128 || K < 32 || N < 64 || (N % 64) != 0) return false;
    if (N > 64 && (N % 128) != 0) return false;
    if (!(pool_k == 1 || pool_k == 32 || pool_k == 64 || (pool_k >= 128 && pool_k % 128 == 0))) return false;
    if (pool_k > 1 && rows % pool_k != 0) return false;
    return true;
}

// Operand split of the inference launches: 2 = fp16x2 with the np = 3 rerun guard (default), 3 = bf16x3 only (psa_set_mlp_mode(2)).
static std::atomic<int> g_tc_np{2};        // process-wide settings: atomics, so that a concurrent psa_set_mlp_mode is a race-free (if unordered) switch
int tc_np() { return g_tc_np; }

// Workspace reservation for one dense layer's images: an fp16x2 image (used when the caller brought no prebuilt one) followed by
// the bf16x3 image of the guarded rerun / of mode 2 / of the training forward.
static size_t tc_dense_image_off3(int K, int N) { return tc_image_alloc_bytes((K + 63) & ~63, N, 2); }
size_t tc_dense_image_bytes(int K, int N) { return tc_dense_image_off3(K, N) + tc_image_alloc_bytes((K + 63) & ~63, N, 3); }

// bytes of a prebuilt image of the current split: mode 0 = fp16x2 blocks + their bf16x3 twin, mode 2 = bf16x3 blocks
static size_t tc_plan_image_bytes(int Kp, int N) { return g_tc_np == 2 ? tc_image_alloc_bytes(Kp, N, 2) + tc_image_alloc_bytes(Kp, N, 3) : tc_image_alloc_bytes(Kp, N, 3); }

// tile width of a dense layer's weight image, with the format flag of the current split
int tc_dense_nt(int N) { return ((N % 128) == 0 ? 128 : 64) | image_flag(g_tc_np); }

// builds the image of W (K x N, rows K..Kp zero) in the format `Nt` carries (width | format flag); zeroes the trailer first
static int build_image(int K, int Kp, int N, int Nt, const float* W, uint8_t* image, cudaStream_t st, const unsigned int* run_if = nullptr) {
    const int np = (Nt & kImageF16x2) ? 2 : 3;
    unsigned int* trailer = reinterpret_cast<unsigned int*>(image + ((tc_image_bytes(Kp, N, np) + 255) & ~(size_t)255));
    if (np == 2) {
        PSA_CUDA(cudaMemsetAsync(trailer, 0, 256, st));
        tc_prep_weights_kernel<2><<<(Kp * N + 255) / 256, 256, 0, st>>>(K, Kp, N, Nt & ~kImageFlags, W, image, trailer, run_if);
    } else {
        tc_prep_weights_kernel<3><<<(Kp * N + 255) / 256, 256, 0, st>>>(K, Kp, N, Nt & ~kImageFlags, W, image, trailer, run_if);
    }
    return check_launch("tc_prep_weights_kernel");
}
static const unsigned int* image_trailer(const uint8_t* image, int Kp, int N, int np) {
    return reinterpret_cast<const unsigned int*>(image + ((tc_image_bytes(Kp, N, np) + 255) & ~(size_t)255));
}

// one launch of a dense layer with NP pieces; `image` holds the weights in that format
template <int NP>
static int launch_tc_dense_np(TcDenseArgs& a, int Nt, cudaStream_t st) {
    dim3 grid2((unsigned)((a.rows + 127) / 128), a.N / Nt);
    const bool big = a.pool_k > 128;
    if (big) { int rc0 = launch_fill_ord_neg_inf(a.rows / a.pool_k * a.N, a.out, st, a.run_if); if (rc0 != PSA_OK) return rc0; }
    if (Nt == 128 && a.Kp <= 512) {
        // transposed kernel, both operands from shared memory; two channel tiles per CTA where the operands are small enough (fp16x2)
        // and the halved grid still covers half of the SMs.  (Pairing always: +0.6 % clouds/s with four batches in flight -- SM-time
        // per output drops -- but SA3 alone 63 -> 76 us, so small grids keep one tile per CTA.)
        const bool pair = NP == 2 && (a.N % 256) == 0 && a.stat_partial == nullptr && (long long)grid2.x * (a.N / 256) >= kNumSMs / 2;
        if (pair) {
            constexpr int CPn = NP == 2 ? 2 : 1;
            const size_t smem3 = (2 * CPn + 2) * (size_t)NP * TcDense3::kPiece + 1024;
            grid2.y = a.N / (128 * CPn);
            PSA_CUDA(cudaFuncSetAttribute(tc_dense3_kernel<NP, CPn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            tc_dense3_kernel<NP, CPn><<<grid2, TcDense3::kThreads, smem3, st>>>(a);
        } else {
            const size_t smem3 = 4 * (size_t)NP * TcDense3::kPiece + 1024;
            PSA_CUDA(cudaFuncSetAttribute(tc_dense3_kernel<NP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            tc_dense3_kernel<NP, 1><<<grid2, TcDense3::kThreads, smem3, st>>>(a);
        }
    } else {
        const size_t smem2 = 4 * (size_t)tc_block_bytes(Nt, NP) + 1024;      // two slots of two 64-K blocks
        if (Nt == 128) {
            PSA_CUDA(cudaFuncSetAttribute(tc_dense2_kernel<128, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            tc_dense2_kernel<128, NP><<<grid2, TcDense2::kThreads, smem2, st>>>(a);
        } else {
            PSA_CUDA(cudaFuncSetAttribute(tc_dense2_kernel<64, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            tc_dense2_kernel<64, NP><<<grid2, TcDense2::kThreads, smem2, st>>>(a);
        }
    }
    int rc = check_launch("tc_dense_kernel");
    if (rc != PSA_OK) return rc;
    if (big) return launch_decode_ord(a.rows / a.pool_k * a.N, a.out, st, a.run_if);
    return PSA_OK;
}

// out = relu?((x . W [+ xyz3 . w3]) * scale + shift) on the tensor cores, optional max over runs of pool_k rows.
//   prebuilt : image of W in the CURRENT split's format (psa_prepare_weight_image), or null
//   ws_img   : tc_dense_image_bytes(K, N) of scratch (missing images are built here)
//   flag     : one zeroed device word (np = 2: raised when a value left the fp16 range; the bf16x3 rerun is conditional on it)
int launch_tc_dense(long long rows, int K, int N, int pool_k, int relu, const float* x, const float* W, const float* scale,
                    const float* shift, float* out, const uint8_t* prebuilt, uint8_t* ws_img, unsigned int* flag, cudaStream_t st,
                    const float* xyz3 = nullptr, const float* w3 = nullptr) {
    const int Kp = (K + 63) & ~63;
    const int Nt = (N % 128) == 0 ? 128 : 64;
    TcDenseArgs a;
    a.rows = rows; a.K = K; a.Kp = Kp; a.N = N; a.pool_k = pool_k; a.relu = relu;
    a.x = x; a.scale = scale; a.shift = shift; a.out = out; a.xyz3 = xyz3; a.w3 = w3;
    uint8_t* img3 = ws_img + tc_dense_image_off3(K, N);
    int rc;
    if (g_tc_np == 3) {
        if (prebuilt == nullptr) { rc = build_image(K, Kp, N, Nt | kImageBf16x3, W, img3, st); if (rc != PSA_OK) return rc; }
        a.image = prebuilt ? prebuilt : img3;
        return launch_tc_dense_np<3>(a, Nt, st);
    }
    if (prebuilt == nullptr) { rc = build_image(K, Kp, N, Nt | kImageF16x2, W, ws_img, st); if (rc != PSA_OK) return rc; }
    a.image = prebuilt ? prebuilt : ws_img;
    a.ovf = flag; a.wflag = image_trailer(a.image, Kp, N, 2);
    rc = launch_tc_dense_np<2>(a, Nt, st);
    if (rc != PSA_OK) return rc;
    // guarded rerun, a no-op unless the fp16x2 pass raised the flag.  A prebuilt image carries its bf16x3 twin behind the fp16x2
    // blocks (psa_prepare_weight_image); otherwise the twin is built here, conditionally too.
    if (prebuilt != nullptr) {
        img3 = const_cast<uint8_t*>(prebuilt) + tc_image_alloc_bytes(Kp, N, 2);
    } else {
        rc = build_image(K, Kp, N, Nt | kImageBf16x3, W, img3, st, flag);
        if (rc != PSA_OK) return rc;
    }
    a.image = img3; a.ovf = nullptr; a.wflag = nullptr; a.run_if = flag;
    return launch_tc_dense_np<3>(a, Nt, st);
}

// Training-mode forward of one layer on tc_dense3: y = relu(bn_prev(x)) . W + bias (pre-BN output), per-row-tile column
// statistics.  The weights change every step, so the image is rebuilt into `image_ws` (tc_dense_image_bytes(K, N)) per call.
// bf16x3 only: batch-statistics activations are not range-checked.
bool tc_train_fwd_eligible(long long rows, int K, int N) {
    // one CTA per 128-row tile with ~14 k cycles of fixed prologue / epilogue: pays off from two K blocks up (K = 64 layers over
    // 500 k rows run faster on the fp32 FMA kernel: 268 vs 302 us measured at SA1's 64 -> 128 layer)
    return rows >= 128 && K >= 128 && K <= 512 && N >= 128 && (N % 128) == 0;
}
int launch_tc_dense_train(long long rows, int K, int N, const float* x, const float* in_scale, const float* in_shift, int in_relu,
                          const float* W, const float* bias, float* y, float* stat_partial, uint8_t* image_ws, cudaStream_t st) {
    const int Kp = (K + 63) & ~63;
    int rc = build_image(K, Kp, N, 128 | kImageBf16x3, W, image_ws, st);
    if (rc != PSA_OK) return rc;
    TcDenseArgs a;
    a.rows = rows; a.K = K; a.Kp = Kp; a.N = N; a.pool_k = 1; a.relu = 0;
    a.x = x; a.image = image_ws; a.scale = nullptr; a.shift = bias; a.out = y; a.xyz3 = nullptr; a.w3 = nullptr;
    a.in_scale = in_scale; a.in_shift = in_shift; a.in_relu = in_relu; a.stat_partial = stat_partial;
    return launch_tc_dense_np<3>(a, 128, st);
}

// A prebuilt image (psa_prepare_weight_image) is used when its format, tile width and first row match; otherwise null (the
// launchers then build what they need in their workspace).  `Nt` carries the format flag.
static const uint8_t* prebuilt_image(const psa_mlp* mlp, int l, int row0, int Nt) {
    if (mlp->image[l] != nullptr && mlp->image_nt[l] == Nt && mlp->image_row0[l] == row0) return reinterpret_cast<const uint8_t*>(mlp->image[l]);
    return nullptr;
}

// PSA_SA_TMODE=0 keeps the last layer of 64-wide levels in row form (A/B runs)
static bool tmode_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PSA_SA_TMODE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

// Can this MLP / geometry run on the tensor-core kernel with `np` pieces per operand?  (otherwise the fp32-FMA fused kernel
// in mlp.cu is used).  Whatever fits with three pieces fits with two; the guarded default needs both.
bool tc_sa_eligible(const psa_mlp* mlp, int c, int nsample, TcArgs* out, int np) {
    if (mlp->n_layers < 2 || mlp->n_layers > 1 + kMaxTcLayers) return false;
    if (!(nsample == 32 || nsample == 64 || nsample == 128)) return false;
    if (mlp->channels[0] != 3 + c) return false;
    const int C1 = mlp->channels[1];
    if (!(C1 == 64 || C1 == 128)) return false;
    TcArgs a{};
    a.C1 = C1;
    a.nl = mlp->n_layers - 1;
    a.np = np;
    for (int l = 0; l < a.nl; ++l) {
        a.Kd[l] = mlp->channels[1 + l];
        a.Ntot[l] = mlp->channels[2 + l];
        if (!(a.Kd[l] == 64 || a.Kd[l] == 128)) return false;
        const bool is_last = (l == a.nl - 1);
        if (!is_last && !(a.Ntot[l] == 64 || a.Ntot[l] == 128)) return false;       // D and the next A operand are one tile wide
        if (is_last && !(a.Ntot[l] == 64 || a.Ntot[l] % 128 == 0)) return false;
    }
    // two row groups per CTA, 64-wide tiles; the last layer is streamed per group if it does not fit; a level that does not fit
    // even then runs on the fp32-FMA fused kernel
    a.dual = 1; a.ntcap = 64; a.stream_last = 0; a.tmode = 0;
    if (tc_dual_layout(a).total + 1024 > 226u * 1024u) a.stream_last = 1;
    if (tc_dual_layout(a).total + 1024 > 226u * 1024u) return false;
    // transposed last layer (lane = channel): 64-wide levels ending in a 128-wide layer, fp16x2 operands, whole neighbourhoods per half tile
    bool t = np == 2 && !a.stream_last && C1 == 64 && a.Ntot[a.nl - 1] == 128 && (nsample == 32 || nsample == 64) && tmode_enabled();
    for (int l = 0; l < a.nl; ++l) t = t && a.Kd[l] == 64;
    if (t) {
        a.tmode = 1;
        if (tc_dual_layout(a).total + 1024 > 226u * 1024u) a.tmode = 0;
    }
    *out = a;
    return true;
}
bool tc_sa_eligible(const psa_mlp* mlp, int c, int nsample, TcArgs* out) {
    TcArgs a3;
    if (!tc_sa_eligible(mlp, c, nsample, &a3, 3)) return false;
    if (g_tc_np == 3) { *out = a3; return true; }
    return tc_sa_eligible(mlp, c, nsample, out, 2);
}

// per tensor layer: an fp16x2 image (when the caller brought none) + the bf16x3 image of the guarded rerun / of mode 2
size_t tc_sa_workspace_bytes(const TcArgs& a, int b, int n, int c) {
    size_t bytes = 256;       // tile counters + range flags
    for (int l = 0; l < a.nl; ++l) bytes += tc_image_alloc_bytes(a.Kd[l], a.Ntot[l], 2) + tc_image_alloc_bytes(a.Kd[l], a.Ntot[l], 3);
    if (c > 0) bytes += (((size_t)b * n * a.C1 * sizeof(float) + 255) & ~(size_t)255) + tc_dense_image_bytes(c, a.C1);
    return bytes;
}

// tile width (with the image-format flag of the current split) of tensor layer l of a level: 64-wide blocks, except the last layer of
// a transposed-mode level, which is the A operand of an M = 128 MMA (one [128 ch][64 k] block)
static int tc_sa_image_nt(const TcArgs& a, int l) { return ((a.tmode && l == a.nl - 1) ? 128 : TcDual::kNt) | image_flag(g_tc_np); }

template <int NP>
static int launch_tc_sa_np(TcArgs& a, cudaStream_t st) {
    const int G = 128 / a.K;
    const long long ntiles = (a.groups + G - 1) / G;
    const size_t smem = (size_t)tc_dual_layout(a).total + 1024;
    long long ctas = (ntiles + 1) / 2;
    if (ctas > kNumSMs) ctas = kNumSMs;
    if (ctas < 1) ctas = 1;
    if (a.tmode) {
        if constexpr (NP == 2) {
            PSA_CUDA(cudaFuncSetAttribute(tc_sa_dual_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc_sa_dual_kernel<3, 2><<<(int)ctas, TcDual::kThreads, smem, st>>>(a);
            return check_launch("tc_sa_dual_kernel");
        }
    }
    const bool pairs = !a.stream_last && (a.Ntot[a.nl - 1] % 128) == 0;            // tile pairs: an even number of resident 64-wide tiles
    bool narrow = a.C1 <= 64;
    for (int l = 0; l < a.nl; ++l) narrow = narrow && a.Kd[l] <= 64;
    if (pairs && narrow) {
        PSA_CUDA(cudaFuncSetAttribute(tc_sa_dual_kernel<1, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_sa_dual_kernel<1, NP><<<(int)ctas, TcDual::kThreads, smem, st>>>(a);
    } else if (pairs && NP == 2) {
        PSA_CUDA(cudaFuncSetAttribute(tc_sa_dual_kernel<(NP == 2 ? 2 : 0), NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_sa_dual_kernel<(NP == 2 ? 2 : 0), NP><<<(int)ctas, TcDual::kThreads, smem, st>>>(a);
    } else {
        PSA_CUDA(cudaFuncSetAttribute(tc_sa_dual_kernel<0, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_sa_dual_kernel<0, NP><<<(int)ctas, TcDual::kThreads, smem, st>>>(a);
    }
    return check_launch("tc_sa_dual_kernel");
}

// One set-abstraction level on the dual-group kernel (`a` = eligibility result for the current split): weight images, the U GEMM
// of the feature part of layer 1, the level, and -- fp16x2 -- its guarded bf16x3 rerun.  w1c: optional centre weights (EdgeConv).
static int tc_sa_run(TcArgs& a, int b, int n, int m, int c, int nsample, const float* xyz, const float* new_xyz, const float* points,
                     const int* idx, const psa_mlp* mlp, const float* w1c, float* out, void* workspace, cudaStream_t st) {
    int rc = PSA_OK;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    unsigned int* words = reinterpret_cast<unsigned int*>(ws);     // [0] tile counter, [1] tile counter of the rerun, [2] range flag of
    PSA_CUDA(cudaMemsetAsync(ws, 0, 256, st));                     // the level, [3] range flag of the U GEMM
    ws += 256;
    uint8_t* img2[kMaxTcLayers];
    uint8_t* img3[kMaxTcLayers];
    for (int l = 0; l < a.nl; ++l) {
        img2[l] = ws; ws += tc_image_alloc_bytes(a.Kd[l], a.Ntot[l], 2);
        img3[l] = ws; ws += tc_image_alloc_bytes(a.Kd[l], a.Ntot[l], 3);
    }
    const float* uf = nullptr;
    if (c > 0) {
        // U = points . W1[3:,:]  once per source point (rows b*n), raw (affine + ReLU are applied after the xyz part)
        float* ufw = reinterpret_cast<float*>(ws);
        uint8_t* uimg = ws + (((size_t)b * n * a.C1 * sizeof(float) + 255) & ~(size_t)255);
        const float* w1f = mlp->weight[0] + (size_t)3 * a.C1;
        if (tc_dense_eligible((long long)b * n, c, a.C1, 1)) {
            rc = launch_tc_dense((long long)b * n, c, a.C1, 1, 0, points, w1f, nullptr, nullptr, ufw, prebuilt_image(mlp, 0, 3, tc_dense_nt(a.C1)), uimg,
                                 words + 3, st);
        } else {
            DenseArgs d;
            d.rows = (long long)b * n; d.K = c; d.N = a.C1; d.pool_k = 1; d.relu = 0;
            d.x = points; d.W = w1f; d.scale = nullptr; d.shift = nullptr; d.out = ufw;
            rc = launch_dense(d, st);
        }
        if (rc != PSA_OK) return rc;
        uf = ufw;
    }
    auto fill = [&](TcArgs& t) {
        t.groups = (long long)b * m; t.K = nsample; t.n = n; t.m = m;
        t.xyz = xyz; t.new_xyz = new_xyz; t.idx = idx; t.out = out; t.uf = uf;
        t.w1x = mlp->weight[0]; t.s1 = mlp->scale[0]; t.t1 = mlp->shift[0]; t.relu1 = mlp->relu[0];
        t.ovf = nullptr; t.run_if = nullptr; t.w1c = w1c;
        for (int l = 0; l < t.nl; ++l) { t.s[l] = mlp->scale[1 + l]; t.t[l] = mlp->shift[1 + l]; t.relu[l] = mlp->relu[1 + l]; t.wflag[l] = nullptr; }
    };
    fill(a);
    for (int l = 0; l < a.nl; ++l) {
        const int nt_img = tc_sa_image_nt(a, l);
        const uint8_t* pre = prebuilt_image(mlp, 1 + l, 0, nt_img);
        uint8_t* own = a.np == 2 ? img2[l] : img3[l];
        if (pre == nullptr) { rc = build_image(a.Kd[l], a.Kd[l], a.Ntot[l], nt_img, mlp->weight[1 + l], own, st); if (rc != PSA_OK) return rc; }
        a.image[l] = pre ? pre : own;
        if (a.np == 2) a.wflag[l] = image_trailer(a.image[l], a.Kd[l], a.Ntot[l], 2);
    }
    a.tile_counter = words;
    if (a.np == 3) return launch_tc_sa_np<3>(a, st);
    a.ovf = words + 2;
    rc = launch_tc_sa_np<2>(a, st);
    if (rc != PSA_OK) return rc;
    // guarded rerun with bf16x3 operands: image builds and the level itself are no-ops unless the fp16x2 pass raised the flag
    TcArgs a3;
    PSA_REQUIRE(tc_sa_eligible(mlp, c, nsample, &a3, 3), "sa_module: internal error (bf16x3 eligibility)");
    fill(a3);
    for (int l = 0; l < a3.nl; ++l) {
        // a prebuilt fp16x2 image carries its bf16x3 twin behind it -- usable here if it has the 64-wide blocks of the row-form kernel
        // (the last layer of a transposed-mode level does not: its twin is rebuilt, conditionally)
        const uint8_t* pre = prebuilt_image(mlp, 1 + l, 0, TcDual::kNt | kImageF16x2);
        if (pre != nullptr) {
            a3.image[l] = pre + tc_image_alloc_bytes(a3.Kd[l], a3.Ntot[l], 2);
        } else {
            rc = build_image(a3.Kd[l], a3.Kd[l], a3.Ntot[l], TcDual::kNt | kImageBf16x3, mlp->weight[1 + l], img3[l], st, words + 2);
            if (rc != PSA_OK) return rc;
            a3.image[l] = img3[l];
        }
    }
    a3.tile_counter = words + 1;
    a3.run_if = words + 2;
    return launch_tc_sa_np<3>(a3, st);
}

}  // namespace psa

using namespace psa;

static std::atomic<int> g_mlp_mode{0};
extern "C" PSA_API int psa_set_mlp_mode(int mode) {
    PSA_REQUIRE(mode == 0 || mode == 1 || mode == 2,
                "set_mlp_mode: mode must be 0 (tensor cores, fp16x2 operands with the range guard), 1 (fp32 FMA kernels only) or 2 (tensor cores, bf16x3)");
    g_mlp_mode = mode;
    g_tc_np = mode == 2 ? 3 : 2;
    return PSA_OK;
}
extern "C" PSA_API int psa_get_mlp_mode(void) { return g_mlp_mode; }

extern "C" size_t psa_sa_module_workspace_bytes(int b, int n, int m, int c, int nsample, const psa_mlp* mlp) {
    (void)m;
    TcArgs a;
    if (g_mlp_mode != 1 && mlp != nullptr && tc_sa_eligible(mlp, c, nsample, &a)) return tc_sa_workspace_bytes(a, b, n, c);
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
    if (g_mlp_mode != 1 && tc_sa_eligible(mlp, c, nsample, &a)) {
        const size_t need = tc_sa_workspace_bytes(a, b, n, c);
        PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need,
                    "sa_module: workspace of %zu bytes required (psa_sa_module_workspace_bytes), got %zu", need, workspace_bytes);
        PSA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sa_module: workspace must be 256-byte aligned");
        return tc_sa_run(a, b, n, m, c, nsample, xyz, new_xyz, points, idx, mlp, nullptr, out, workspace, st);
    }
    return sa_module_simt(b, n, m, c, nsample, xyz, new_xyz, points, idx, mlp, out, st);
}

extern "C" size_t psa_shared_mlp_workspace_bytes(long long rows, const psa_mlp* mlp) {
    if (mlp == nullptr || mlp->n_layers < 1) return 0;
    size_t bytes = 0;
    int cmax = 0;
    for (int l = 1; l < mlp->n_layers; ++l) cmax = cmax > mlp->channels[l] ? cmax : mlp->channels[l];
    if (mlp->n_layers > 1) bytes += 2 * (((size_t)rows * cmax * sizeof(float) + 255) & ~(size_t)255);
    for (int l = 0; l < mlp->n_layers; ++l) bytes += tc_dense_image_bytes(mlp->channels[l], mlp->channels[l + 1] < 64 ? 64 : mlp->channels[l + 1]);
    if (rows <= 32) {
        size_t fc = 0;
        for (int l = 0; l < mlp->n_layers; ++l) { size_t f = fc_small_workspace_bytes(mlp->channels[l], mlp->channels[l + 1]); fc = f > fc ? f : fc; }
        bytes += (fc + 255) & ~(size_t)255;
    }
    return bytes + 256;       // range flags of the tensor-core layers (one word per layer), last
}

extern "C" int psa_shared_mlp(long long rows, int pool_k, const float* x, const psa_mlp* mlp, float* out,
                              void* workspace, size_t workspace_bytes, psa_stream_t stream) {
    int rc = validate_mlp_public(mlp, "shared_mlp");
    if (rc != PSA_OK) return rc;
    PSA_REQUIRE(rows >= 0 && pool_k >= 1, "shared_mlp: rows=%lld pool_k=%d", rows, pool_k);
    if (rows == 0) return PSA_OK;
    PSA_REQUIRE(rows % pool_k == 0, "shared_mlp: rows=%lld is not a multiple of pool_k=%d", rows, pool_k);
    PSA_REQUIRE(x && out, "shared_mlp: null buffer");
    const int L = mlp->n_layers;
    const size_t need = psa_shared_mlp_workspace_bytes(rows, mlp);
    PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, "shared_mlp: workspace of %zu bytes required (got %zu)", need, workspace_bytes);
    PSA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "shared_mlp: workspace must be 256-byte aligned");
    int cmax = 0;
    for (int l = 1; l < L; ++l) cmax = cmax > mlp->channels[l] ? cmax : mlp->channels[l];
    const size_t half = L > 1 ? (((size_t)rows * cmax * sizeof(float) + 255) & ~(size_t)255) : 0;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
    float* ws0 = reinterpret_cast<float*>(wsb);
    float* ws1 = reinterpret_cast<float*>(wsb + half);
    uint8_t* img = wsb + 2 * half;
    float* fc_partial = nullptr;
    if (rows <= 32) {   // the K-split partial sums of the small-M kernel sit at the end of the workspace
        size_t imgs = 0;
        for (int l = 0; l < L; ++l) imgs += tc_dense_image_bytes(mlp->channels[l], mlp->channels[l + 1] < 64 ? 64 : mlp->channels[l + 1]);
        fc_partial = reinterpret_cast<float*>(img + imgs);
    }
    const float* cur = x;
    cudaStream_t st = as_stream(stream);
    unsigned int* flags = reinterpret_cast<unsigned int*>(wsb + need - 256);
    PSA_CUDA(cudaMemsetAsync(flags, 0, 256, st));
    for (int l = 0; l < L; ++l) {
        const int K = mlp->channels[l], N = mlp->channels[l + 1];
        const int pk = (l == L - 1) ? pool_k : 1;
        float* dst = (l == L - 1) ? out : ((l & 1) ? ws1 : ws0);
        if (g_mlp_mode != 1 && tc_dense_eligible(rows, K, N, pk)) {
            rc = launch_tc_dense(rows, K, N, pk, mlp->relu[l], cur, mlp->weight[l], mlp->scale[l], mlp->shift[l], dst, prebuilt_image(mlp, l, 0, tc_dense_nt(N)),
                                 img, flags + l, st);
        } else {
            DenseArgs d;
            d.rows = rows; d.K = K; d.N = N; d.pool_k = pk; d.relu = mlp->relu[l];
            d.x = cur; d.W = mlp->weight[l]; d.scale = mlp->scale[l]; d.shift = mlp->shift[l]; d.out = dst;
            rc = (rows <= 32 && pk == 1) ? launch_fc_small(d, fc_partial, st) : launch_dense(d, st);
        }
        if (rc != PSA_OK) return rc;
        img += tc_dense_image_bytes(K, N < 64 ? 64 : N);
        cur = dst;
    }
    return PSA_OK;
}

// pointnet_sa_module(group_all=True) (pointnet_util.py:59-84,113-127): rows = [xyz, points] (xyz first), MLP, max over the
// n points of each cloud -- without building the (b,n,3+c) concatenation: the feature part [3:,:] of the first layer runs
// as an aligned K = c GEMM on the tensor cores, the three xyz rows of W1 are folded into its epilogue.
extern "C" size_t psa_sa_group_all_workspace_bytes(int b, int n, int c, const psa_mlp* mlp) {
    if (mlp == nullptr || mlp->n_layers < 1) return 0;
    psa_mlp m2 = *mlp;
    m2.channels[0] = c;
    return psa_shared_mlp_workspace_bytes((long long)b * n, &m2);
}

extern "C" int psa_sa_group_all_infer(int b, int n, int c, const float* xyz, const float* points, const psa_mlp* mlp,
                                      float* out, void* workspace, size_t workspace_bytes, psa_stream_t stream) {
    int rc = validate_mlp_public(mlp, "sa_group_all");
    if (rc != PSA_OK) return rc;
    PSA_REQUIRE(b >= 0 && n >= 1 && c >= 1, "sa_group_all: bad dims b=%d n=%d c=%d", b, n, c);
    PSA_REQUIRE(mlp->channels[0] == 3 + c, "sa_group_all: mlp input width %d != 3 + c (%d)", mlp->channels[0], 3 + c);
    if (b == 0) return PSA_OK;
    PSA_REQUIRE(xyz && points && out, "sa_group_all: null buffer");
    const long long rows = (long long)b * n;
    const int L = mlp->n_layers;
    const int N0 = mlp->channels[1];
    const int pk0 = (L == 1) ? n : 1;
    PSA_SUPPORTED(g_mlp_mode != 1 && tc_dense_eligible(rows, c, N0, pk0),
                  "sa_group_all: first layer (%d -> %d over %lld rows) is not eligible for the tensor-core path; concatenate and use shared_mlp", c, N0, rows);
    const size_t need = psa_sa_group_all_workspace_bytes(b, n, c, mlp);
    PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, "sa_group_all: workspace of %zu bytes required (got %zu)", need, workspace_bytes);
    PSA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sa_group_all: workspace must be 256-byte aligned");
    int cmax = 0;
    for (int l = 1; l < L; ++l) cmax = cmax > mlp->channels[l] ? cmax : mlp->channels[l];
    const size_t half = L > 1 ? (((size_t)rows * cmax * sizeof(float) + 255) & ~(size_t)255) : 0;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
    float* ws0 = reinterpret_cast<float*>(wsb);
    float* ws1 = reinterpret_cast<float*>(wsb + half);
    uint8_t* img = wsb + 2 * half;
    cudaStream_t st = as_stream(stream);
    unsigned int* flags = reinterpret_cast<unsigned int*>(wsb + need - 256);
    PSA_CUDA(cudaMemsetAsync(flags, 0, 256, st));
    const float* cur = points;
    for (int l = 0; l < L; ++l) {
        const int K = (l == 0) ? c : mlp->channels[l], N = mlp->channels[l + 1];
        const int pk = (l == L - 1) ? n : 1;
        float* dst = (l == L - 1) ? out : ((l & 1) ? ws1 : ws0);
        const float* W = (l == 0) ? mlp->weight[0] + (size_t)3 * N : mlp->weight[l];
        if (tc_dense_eligible(rows, K, N, pk)) {
            rc = launch_tc_dense(rows, K, N, pk, mlp->relu[l], cur, W, mlp->scale[l], mlp->shift[l], dst, prebuilt_image(mlp, l, l == 0 ? 3 : 0, tc_dense_nt(N)),
                                 img, flags + l, st, l == 0 ? xyz : nullptr, l == 0 ? mlp->weight[0] : nullptr);
        } else {
            PSA_SUPPORTED(l > 0, "sa_group_all: layer 0 must run on the tensor-core path");
            DenseArgs d;
            d.rows = rows; d.K = K; d.N = N; d.pool_k = pk; d.relu = mlp->relu[l];
            d.x = cur; d.W = W; d.scale = mlp->scale[l]; d.shift = mlp->shift[l]; d.out = dst;
            rc = launch_dense(d, st);
        }
        if (rc != PSA_OK) return rc;
        img += tc_dense_image_bytes(K, N < 64 ? 64 : N);
        cur = dst;
    }
    return PSA_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// EdgeConv, single layer (dgcnn/models/dgcnn.py:41-47 etc.):  max_j relu(BN(W . [x_i ; x_j - x_i] + b)).
//   W . [x_i ; x_j - x_i] = (W_a - W_b) . x_i + W_b . x_j   and, per channel, relu(s*(A_i + B_j) + t) is monotone in B_j
//   (increasing if s >= 0, decreasing otherwise), so the max over the k edges only needs max_j / min_j of B:
//   one (B*N) x C x 2C_out GEMM over POINTS (k-fold fewer rows than the reference's conv over B*N*k edges, tensor cores)
//   + one gather-max pass.  No (B,N,k,2C) edge tensor, no (B,N,k,C_out) activation tensor.
// ------------------------------------------------------------------------------------------------------------------
__global__ void edge_wc_kernel(int c, int N, const float* __restrict__ W, float* __restrict__ Wc) {
    const int total = c * 2 * N;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kk = e / (2 * N), j = e - kk * 2 * N;
        Wc[e] = j < N ? __ldg(W + (size_t)kk * N + j) - __ldg(W + (size_t)(c + kk) * N + j) : __ldg(W + (size_t)(c + kk) * N + j - N);
    }
}

template <int VEC>   // channels per lane: N = 32 * VEC
__global__ void __launch_bounds__(256)
edge_gather_max_kernel(long long points, int n, int k, int N, const float* __restrict__ AB, const int* __restrict__ nn_idx,
                       const float* __restrict__ scale, const float* __restrict__ shift, int relu, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        sc[v] = scale ? __ldg(scale + lane * VEC + v) : 1.f;
        sh[v] = shift ? __ldg(shift + lane * VEC + v) : 0.f;
    }
    for (long long p = warp0; p < points; p += (long long)gridDim.x * 8) {
        const long long base = (p / n) * n;
        const float* arow = AB + (size_t)p * 2 * N + lane * VEC;
        float a[VEC], mx[VEC], mn[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { a[v] = __ldg(arow + v); mx[v] = -FLT_MAX; mn[v] = FLT_MAX; }
        const int myj = lane < k ? __ldg(nn_idx + p * k + lane) : 0;          // k <= 32
        for (int j = 0; j < k; ++j) {
            const int nb = __shfl_sync(0xffffffffu, myj, j);
            const float* brow = AB + (size_t)(base + nb) * 2 * N + N + lane * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { const float bv = __ldg(brow + v); mx[v] = fmaxf(mx[v], bv); mn[v] = fminf(mn[v], bv); }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float y = fmaf(a[v] + (sc[v] >= 0.f ? mx[v] : mn[v]), sc[v], sh[v]);
            if (relu) y = fmaxf(y, 0.f);
            out[(size_t)p * N + lane * VEC + v] = y;
        }
    }
}

static bool edgeconv_algebra_ok(long long rows, int c, int k, const psa_mlp* mlp) {
    const int N = mlp->channels[1];
    return g_mlp_mode != 1 && mlp->n_layers == 1 && rows >= 128 && k <= 32 && (N == 32 || N == 64 || N == 128 || N == 256) && c >= 1;
}

// Multi-layer EdgeConv over 3-D points (DGCNN's input transform net, dgcnn/models/transform_nets.py:13-27: [x_i, x_j - x_i] ->
// 64 -> 128 -> max over k) on the set-abstraction kernel: W1 . [x_i ; x_j - x_i] = W1[0:3] . x_i + W1[3:6] . (x_j - x_i) is exactly
// a level whose centres are the points themselves, with W1[3:6] as the xyz weights and W1[0:3] as centre weights (TcArgs::w1c); the
// k <= 32 neighbours are padded to the kernel's 32-row neighbourhoods by repeating the first one (a max-pool ignores duplicates).
__global__ void edge_pad_idx_kernel(long long rows, int k, const int* __restrict__ idx, int* __restrict__ idx32) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < rows * 32; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e >> 5;
        const int j = (int)(e & 31);
        idx32[e] = __ldg(idx + r * k + (j < k ? j : 0));
    }
}
static bool edgeconv_dual_ok(int c, int k, const psa_mlp* mlp, psa_mlp* m2, TcArgs* a) {
    if (g_mlp_mode == 1 || c != 3 || k < 1 || k > 32 || mlp->n_layers < 2) return false;
    *m2 = *mlp;
    m2->channels[0] = 3;
    m2->weight[0] = mlp->weight[0] + (size_t)3 * mlp->channels[1];
    for (int l = 0; l < PSA_MAX_MLP_LAYERS; ++l) { m2->image[l] = nullptr; m2->image_nt[l] = 0; m2->image_row0[l] = 0; }
    return tc_sa_eligible(m2, 0, 32, a);
}

extern "C" size_t psa_edgeconv_workspace_bytes(int b, int n, int c, int k, const psa_mlp* mlp) {
    if (mlp == nullptr) return 0;
    const long long rows = (long long)b * n;
    {
        psa_mlp m2;
        TcArgs a;
        if (edgeconv_dual_ok(c, k, mlp, &m2, &a)) return (((size_t)rows * 32 * 4 + 255) & ~(size_t)255) + tc_sa_workspace_bytes(a, b, n, 0);
    }
    if (!edgeconv_algebra_ok(rows, c, k, mlp)) return 0;
    const int N = mlp->channels[1];
    return (((size_t)c * 2 * N * 4 + 255) & ~(size_t)255) + (((size_t)rows * 2 * N * 4 + 255) & ~(size_t)255) + tc_dense_image_bytes(c, 2 * N) + 256;
}

extern "C" int psa_edgeconv_infer(int b, int n, int c, int k, const float* x, const int* nn_idx, const psa_mlp* mlp,
                                  float* out, void* workspace, size_t workspace_bytes, psa_stream_t stream) {
    int rc = validate_mlp_public(mlp, "edgeconv");
    if (rc != PSA_OK) return rc;
    PSA_REQUIRE(b >= 0 && n >= 0 && c >= 1 && k >= 1, "edgeconv: bad dims b=%d n=%d c=%d k=%d", b, n, c, k);
    PSA_REQUIRE(mlp->channels[0] == 2 * c, "edgeconv: mlp input width %d != 2*c (%d)", mlp->channels[0], 2 * c);
    if (b == 0 || n == 0) return PSA_OK;
    PSA_REQUIRE(x && nn_idx && out, "edgeconv: null buffer");
    cudaStream_t st = as_stream(stream);
    const long long rows = (long long)b * n;
    {
        psa_mlp m2;
        TcArgs a;
        if (edgeconv_dual_ok(c, k, mlp, &m2, &a)) {
            const size_t need = psa_edgeconv_workspace_bytes(b, n, c, k, mlp);
            PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, "edgeconv: workspace of %zu bytes required (got %zu)", need, workspace_bytes);
            PSA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "edgeconv: workspace must be 256-byte aligned");
            int* idx32 = reinterpret_cast<int*>(workspace);
            edge_pad_idx_kernel<<<(unsigned)((rows * 32 + 255) / 256 < 65535 * 4 ? (rows * 32 + 255) / 256 : 65535 * 4), 256, 0, st>>>(rows, k, nn_idx, idx32);
            rc = check_launch("edge_pad_idx_kernel");
            if (rc != PSA_OK) return rc;
            return tc_sa_run(a, b, n, n, 0, 32, x, x, nullptr, idx32, &m2, mlp->weight[0], out,
                             reinterpret_cast<uint8_t*>(workspace) + (((size_t)rows * 32 * 4 + 255) & ~(size_t)255), st);
        }
    }
    if (!edgeconv_algebra_ok(rows, c, k, mlp)) return edgeconv_simt(b, n, c, k, x, nn_idx, mlp, out, st);
    const int N = mlp->channels[1];
    const size_t need = psa_edgeconv_workspace_bytes(b, n, c, k, mlp);
    PSA_REQUIRE(workspace != nullptr && workspace_bytes >= need, "edgeconv: workspace of %zu bytes required (got %zu)", need, workspace_bytes);
    PSA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "edgeconv: workspace must be 256-byte aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* Wc = reinterpret_cast<float*>(ws);
    ws += ((size_t)c * 2 * N * 4 + 255) & ~(size_t)255;
    float* AB = reinterpret_cast<float*>(ws);
    ws += ((size_t)rows * 2 * N * 4 + 255) & ~(size_t)255;
    edge_wc_kernel<<<(c * 2 * N + 255) / 256, 256, 0, st>>>(c, N, mlp->weight[0], Wc);
    if (tc_dense_eligible(rows, c, 2 * N, 1)) {
        unsigned int* flag = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(workspace) + need - 256);
        PSA_CUDA(cudaMemsetAsync(flag, 0, 256, st));
        rc = launch_tc_dense(rows, c, 2 * N, 1, 0, x, Wc, nullptr, nullptr, AB, nullptr, ws, flag, st);
    } else {
        DenseArgs d;
        d.rows = rows; d.K = c; d.N = 2 * N; d.pool_k = 1; d.relu = 0;
        d.x = x; d.W = Wc; d.scale = nullptr; d.shift = nullptr; d.out = AB;
        rc = launch_dense(d, st);
    }
    if (rc != PSA_OK) return rc;
    const int grid = (int)((rows + 7) / 8 < (long long)kNumSMs * 8 ? (rows + 7) / 8 : (long long)kNumSMs * 8);
    const int vec = N / 32;
#define PSA_EDGE_LAUNCH(V) edge_gather_max_kernel<V><<<grid, 256, 0, st>>>(rows, n, k, N, AB, nn_idx, mlp->scale[0], mlp->shift[0], mlp->relu[0], out)
    if (vec == 1) PSA_EDGE_LAUNCH(1); else if (vec == 2) PSA_EDGE_LAUNCH(2); else if (vec == 4) PSA_EDGE_LAUNCH(4); else PSA_EDGE_LAUNCH(8);
#undef PSA_EDGE_LAUNCH
    return check_launch("edge_gather_max_kernel");
}

extern "C" int psa_prepare_weight_image(int K, int N, int row0, int nt, const float* W, void* image, psa_stream_t stream) {
    const int ntw = nt & ~kImageFlags;        // nt as returned by psa_mlp_image_plan: tile width | format flag
    PSA_REQUIRE((nt & kImageFlags) == kImageBf16x3 || (nt & kImageFlags) == kImageF16x2, "prepare_weight_image: nt=%d carries no image format flag", nt);
    PSA_REQUIRE(K >= 1 && N >= 64 && N % 64 == 0 && row0 >= 0 && row0 < K && (ntw == 64 || ntw == 128) && N % ntw == 0,
                "prepare_weight_image: bad arguments K=%d N=%d row0=%d nt=%d", K, N, row0, nt);
    PSA_REQUIRE(W && image, "prepare_weight_image: null buffer");
    const int Ki = K - row0, Kp = (Ki + 63) & ~63;
    int rc = build_image(Ki, Kp, N, nt, W + (size_t)row0 * N, reinterpret_cast<uint8_t*>(image), as_stream(stream));
    if (rc != PSA_OK || (nt & kImageF16x2) == 0) return rc;
    // fp16x2 images carry the bf16x3 image of the range guard's rerun right behind them
    return build_image(Ki, Kp, N, ntw | kImageBf16x3, W + (size_t)row0 * N, reinterpret_cast<uint8_t*>(image) + tc_image_alloc_bytes(Kp, N, 2), as_stream(stream));
}

extern "C" int psa_mlp_image_plan(int usage, long long rows, int pool_k, int c, int nsample, const psa_mlp* mlp,
                                  int nt[PSA_MAX_MLP_LAYERS], int row0[PSA_MAX_MLP_LAYERS], size_t bytes[PSA_MAX_MLP_LAYERS]) {
    int rc = validate_mlp_public(mlp, "mlp_image_plan");
    if (rc != PSA_OK) return rc;
    for (int l = 0; l < PSA_MAX_MLP_LAYERS; ++l) { nt[l] = 0; row0[l] = 0; bytes[l] = 0; }
    if (g_mlp_mode == 1) return PSA_OK;
    const int L = mlp->n_layers;
    if (usage == PSA_USAGE_SHARED_MLP || usage == PSA_USAGE_SA_GROUP_ALL) {
        for (int l = 0; l < L; ++l) {
            const int r0 = (usage == PSA_USAGE_SA_GROUP_ALL && l == 0) ? 3 : 0;
            const int K = mlp->channels[l] - r0, N = mlp->channels[l + 1];
            const int pk = (l == L - 1) ? pool_k : 1;
            if (K >= 1 && tc_dense_eligible(rows, K, N, pk)) { nt[l] = tc_dense_nt(N); row0[l] = r0; bytes[l] = tc_plan_image_bytes((K + 63) & ~63, N); }
        }
        return PSA_OK;
    }
    PSA_REQUIRE(usage == PSA_USAGE_SA_MODULE, "mlp_image_plan: unknown usage %d", usage);
    TcArgs a;
    if (!tc_sa_eligible(mlp, c, nsample, &a)) return PSA_OK;
    if (c > 0 && tc_dense_eligible(rows, c, a.C1, 1)) { nt[0] = tc_dense_nt(a.C1); row0[0] = 3; bytes[0] = tc_plan_image_bytes((c + 63) & ~63, a.C1); }
    for (int l = 0; l < a.nl; ++l) { nt[1 + l] = tc_sa_image_nt(a, l); row0[1 + l] = 0; bytes[1 + l] = tc_plan_image_bytes(a.Kd[l], a.Ntot[l]); }
    return PSA_OK;
}

#ifdef PSA_TC_TIMING
extern "C" __attribute__((visibility("default"))) int psa_debug_tc_timing(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    if (out8) cudaMemcpyFromSymbol(out8, psa::g_tc_timing, 8 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(psa::g_tc_timing, z, sizeof(z)); }
    return 0;
}
#endif

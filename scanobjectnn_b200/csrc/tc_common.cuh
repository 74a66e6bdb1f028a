// tc_common.cuh -- hand-written tcgen05 / TMEM / mbarrier wrappers for sm_100a (inline PTX, no CUTLASS).
//
// Conventions used by the kernels in tc_mlp.cu:
//   * accumulators D and the A operand live in TMEM (128 lanes x 512 32-bit columns per SM); row r of a 128-row tile
//     is TMEM lane r, warp w of the CTA's four "row" warps owns lanes 32w..32w+31;
//   * the B operand (weights, [N][K] "K-major") lives in shared memory in the canonical 128-byte-swizzle layout:
//     8-row x 128-byte atoms, 16-byte chunk index XOR (row % 8), atoms of consecutive 8-row groups 1024 B apart (SBO),
//     consecutive 128-byte K blocks N*128 B apart;
//   * one warp runs the issue code converged, its elected lane issues tcgen05.mma and commits to an mbarrier (see MMA
//     ISSUE CONVENTION below); everybody else waits on the barrier's parity.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace psa {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one full warp executes these) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma reads B through descriptors)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
#if defined(PSA_MBAR_SLEEP_NS) && PSA_MBAR_SLEEP_NS > 0
        if (!done) __nanosleep(PSA_MBAR_SLEEP_NS);      // A/B builds (tools/build_variant.py): back off instead of re-polling at once
#endif
    } while (!done);
}
// MMA ISSUE CONVENTION.  tcgen05.mma takes its operands from UNIFORM registers.  Issued from `if (tid == 0)` the
// compiler cannot prove uniformity and wraps every MMA in an ELECT / R2UR.BROADCAST loop (~17 instructions, ~110 cycles
// per MMA measured -- slower than the 32 / 64 cycles the tensor pipe needs for an N = 64 / 128 instruction,
// tools/microbench/mma_rate.cu).  So: the WHOLE issuer warp runs the issue code converged, with operands derived from
// warp-uniform values (kernel parameters, loop counters, `warp_uniform(...)`), and elect.sync picks the lane that
// executes the instruction.  The elected lane is the same every time (lowest active), so tcgen05.commit tracks its MMAs.
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// all previously issued tcgen05.mma of the elected lane arrive on `bar` when they complete (converged warp)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
        : "memory");
}

// ---- bulk async copy global -> shared (TMA engine, 1-D, no tensor map), completion on an mbarrier ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- descriptors ----
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1) |
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout type = 2
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// The same descriptor as base + byte offset: the start-address field is (address >> 4), so moving the tile by `off`
// bytes (a multiple of 16, same 256 KB window) is one add on the low word -- per-MMA descriptor math stays uniform.
struct SmemDescBase { uint32_t lo, hi; };
__device__ __forceinline__ SmemDescBase smem_desc_base(uint32_t smem_addr) {
    SmemDescBase b;
    b.lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
    b.hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return b;
}
__device__ __forceinline__ uint64_t smem_desc_at(SmemDescBase b, uint32_t off) { return ((uint64_t)b.hi << 32) | (uint64_t)(b.lo + (off >> 4)); }
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A/B format (2 = tf32, 1 = bf16), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(uint32_t ab_format, uint32_t M, uint32_t N) {
    return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2;     // kind::f16 takes F16 or BF16 operands; kind::tf32 takes TF32

// D[tmem] (+)= A[tmem] * B[smem desc]; accumulate = 0 overwrites D.  Call from a CONVERGED warp (see above).
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc] (both operands from shared memory).  Converged warp, as above.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- TMEM <-> registers: each lane moves its own row (TMEM lane = 32*(warp%4) + laneid) ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// ---- operand quantisation (done by us, so the tensor core only ever sees exactly representable values) ----
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x = lo_elem (low 16 bits), .y = hi_elem
    return *reinterpret_cast<uint32_t*>(&p);
}

// fp16 pieces (two per operand: 22 mantissa bits, three MMAs per product).  pack: .x = lo_elem in the low 16 bits.
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
    __half2 p = __floats2half2_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t p) { return __half22float2(*reinterpret_cast<__half2*>(&p)); }
// running |max| of every leading piece a kernel stores (packed halves): an infinity here means a value left the fp16 range
__device__ __forceinline__ void track_f16x2(uint32_t& m, uint32_t p) {
    const __half2 r = __hmax2(*reinterpret_cast<__half2*>(&m), __habs2(*reinterpret_cast<__half2*>(&p)));
    m = *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ bool f16x2_overflowed(uint32_t m) { return (m & 0x7c00u) == 0x7c00u || (m & 0x7c000000u) == 0x7c000000u; }

// Operand split of the tensor-core kernels, NP pieces per operand:
//   NP = 3: three bf16 pieces (a = a1 + a2 + a3 exactly), six MMAs  a1w3 + a2w2 + a3w1 + a1w2 + a2w1 + a1w1  (small products first: the
//           accumulator add truncates) -- any fp32 magnitude;
//   NP = 2: two fp16 pieces (22 mantissa bits; the tail below 2^-24 absolute is dropped), three MMAs  a1w2 + a2w1 + a1w1 -- half
//           the tensor work, same accuracy against fp64 as the fp32 FMA kernels, valid while |a| < 65504 (the kernels track the
//           stored pieces and raise a flag otherwise; the launcher then reruns the op on the NP = 3 instantiation).
template <int NP> struct Split;
template <> struct Split<3> {
    static constexpr uint32_t kFmt = kFmtBF16;
    static constexpr int kTerms = 6;
    __host__ __device__ static constexpr uint32_t a(int t) { return t == 0 ? 0u : t == 1 ? 1u : t == 2 ? 2u : t == 3 ? 0u : t == 4 ? 1u : 0u; }
    __host__ __device__ static constexpr uint32_t w(int t) { return t == 0 ? 2u : t == 1 ? 1u : t == 2 ? 0u : t == 3 ? 1u : 0u; }
};
template <> struct Split<2> {
    static constexpr uint32_t kFmt = kFmtF16;
    static constexpr int kTerms = 3;
    __host__ __device__ static constexpr uint32_t a(int t) { return t == 1 ? 1u : 0u; }
    __host__ __device__ static constexpr uint32_t w(int t) { return t == 0 ? 1u : 0u; }
};
// split two consecutive elements into NP packed pieces p[0..NP) (h0, h1 are clobbered); NP = 2 also tracks the leading piece
template <int NP>
__device__ __forceinline__ void split_pair(float h0, float h1, uint32_t (&p)[NP], uint32_t& ovf) {
    if constexpr (NP == 3) {
        (void)ovf;
        p[0] = pack_bf16x2(h0, h1);
        h0 -= __uint_as_float(p[0] << 16); h1 -= __uint_as_float(p[0] & 0xffff0000u);
        p[1] = pack_bf16x2(h0, h1);
        h0 -= __uint_as_float(p[1] << 16); h1 -= __uint_as_float(p[1] & 0xffff0000u);
        p[2] = pack_bf16x2(h0, h1);
    } else {
        p[0] = pack_f16x2(h0, h1);
        track_f16x2(ovf, p[0]);
        const float2 f = unpack_f16x2(p[0]);
        p[1] = pack_f16x2(h0 - f.x, h1 - f.y);
    }
}

// byte offset of element (n, k) of a [N][K] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz_off_f32(uint32_t n, uint32_t k, uint32_t N) {
    return (k >> 5) * (N * 128u) + (n >> 3) * 1024u + (n & 7u) * 128u + ((((k & 31u) >> 2) ^ (n & 7u)) << 4) + (k & 3u) * 4u;
}
__device__ __forceinline__ uint32_t swz_off_bf16(uint32_t n, uint32_t k, uint32_t N) {
    return (k >> 6) * (N * 128u) + (n >> 3) * 1024u + (n & 7u) * 128u + ((((k & 63u) >> 3) ^ (n & 7u)) << 4) + (k & 7u) * 2u;
}

}  // namespace tc
}  // namespace psa

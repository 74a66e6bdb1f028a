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
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------------------------------------------
// Streaming F1 kernel (round 2): the same front as above, organised so that the only thing the SMs do for most of the
// launch is stream the (B,m,K,C1) tensor out.
//   * persistent CTAs (2 per SM), each owning an equal contiguous range of the B*m queries (ranges may cross clouds; the
//     grid of a cloud is rebuilt at a crossing);
//   * warp 0 = PRODUCER: lane-per-slab grid search into per-query bitmaps (ball_query.cuh), bitmap -> ordered idx rows,
//     idx/pts_cnt to global memory, and the centred neighbour coordinates (dx,dy,dz,j) of every row of the batch into a
//     shared-memory ring slot -- ~150 warp instructions per query instead of ~600;
//   * warps 1..7 = CONSUMERS: per step 4 rows x C1 channels -- one LDS.128 for the row's (dx,dy,dz,j), 6 FFMA2 per 4
//     channels with the weights resident in registers, 128-byte streaming stores, BN statistics in registers;
//   * ring of kF1Ring batches of kF1Batch queries, named barriers FULL/EMPTY per slot: the search of batch t+2 runs
//     under the stores of batch t.
// Statistics: per-CTA partials in a fixed order, the last CTA to finish (atomic ticket) adds them in fp64 in CTA order:
// deterministic, one launch (+ a 4-byte memset node for the ticket).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kF1SThreads = 256;
constexpr int kF1SConsWarps = 7;
constexpr int kF1Batch = 8;
constexpr int kF1Ring = 3;

struct F1SArgs {
    int b, n, m, nsample, C1;
    float radius, thr;
    int none, want_grid;
    const float* xyz;
    const float* new_xyz;
    const float* uf;
    const float* w1;
    const float* bias;
    float* pre;
    int* idx;
    int* pts_cnt;
    float* partial;        // (gridDim.x, 2, C1)
    float* stats;          // (2, C1)
    unsigned* ticket;      // zeroed by the launcher
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__host__ __device__ inline size_t f1s_slot_bytes(int nsample) { return (size_t)kF1Batch * nsample * (sizeof(int) + sizeof(float4)); }
// grid mode: cell-sorted cloud + cell table + ONE warp's bitmaps (the producer's; ball_query.cuh lays a per-warp scratch
// area out right behind the cell table); scan mode: the SoA copy
__host__ __device__ inline size_t f1s_bq_bytes(int n, bool grid) {
    const size_t b = grid ? (size_t)n * 16 + (size_t)(kBqMaxCells + 32) * 4 + (size_t)bq_warp_scratch_words(n) * 4 : bq_smem_bytes(n, false);
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t f1s_smem_bytes(int n, int nsample, bool grid) { return f1s_bq_bytes(n, grid) + kF1Ring * f1s_slot_bytes(nsample); }

template <int NV, bool STATS, bool HAS_U, int PPT>
__global__ void __launch_bounds__(kF1SThreads, 2)
sa_conv1_stream_kernel(const __grid_constant__ F1SArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    const int n = a.n, K = a.nsample, C1 = a.C1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_grid = a.want_grid != 0;
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem_f);
    const size_t bq_bytes = f1s_bq_bytes(n, want_grid);
    const int wpl = bq_bitmap_words_per_lane(n);
    unsigned* bitmaps = reinterpret_cast<unsigned*>(bq_carve(smem_f, n, want_grid, nullptr).hits);   // kF1Batch x (32*wpl) words (grid mode)
    uint8_t* ring = sm + bq_bytes;
    const size_t slot_bytes = f1s_slot_bytes(K);
    static_assert(kF1Batch == kBqSlabQueries, "the producer's bitmaps are one warp's scratch area");
    if (want_grid)
        for (int i = tid; i < kF1Batch * 32 * wpl; i += kF1SThreads) bitmaps[i] = 0u;

    // ---- consumer registers: lane-in-row `sub` owns channels [32 i + 4 sub, +4), i < NV ----
    const int sub = lane & 7, rsub = lane >> 3;
    constexpr int NPK = 2 * NV;
    float2 ssum[NPK], ssq[NPK];
#pragma unroll
    for (int p = 0; p < NPK; ++p) ssum[p] = ssq[p] = make_float2(0.f, 0.f);

    const long long T = (long long)a.b * a.m;
    const long long q_begin = T * blockIdx.x / gridDim.x, q_end = T * (blockIdx.x + 1) / gridDim.x;
    // batches this CTA will run in total (the consumers' last kF1Ring EMPTY arrivals have no taker and are skipped)
    int total_batches = 0;
    for (long long q = q_begin; q < q_end;) {
        const long long cloud = q / a.m;
        const long long seg_end = min(q_end, (cloud + 1) * (long long)a.m);
        total_batches += (int)((seg_end - q + kF1Batch - 1) / kF1Batch);
        q = seg_end;
    }
    int ring_pos = 0;

    for (long long q = q_begin; q < q_end;) {
        const long long cloud = q / a.m;
        const long long seg_end = min(q_end, (cloud + 1) * (long long)a.m);
        const int nseg = (int)(seg_end - q);
        const int nb = (nseg + kF1Batch - 1) / kF1Batch;
        const float* gx = a.xyz + (size_t)cloud * n * 3;
        __syncthreads();                                   // the previous cloud's grid is no longer read
        const BqSmem s = bq_carve(smem_f, n, want_grid, gx);
        const BqGrid g = bq_stage_and_build<PPT>(s, n, a.radius, want_grid);

        if (warp == 0) {
            // =========================== PRODUCER ===========================
            for (int bi = 0; bi < nb; ++bi, ++ring_pos) {
                const int slot = ring_pos % kF1Ring;
                if (ring_pos >= kF1Ring) named_bar_sync(1 + kF1Ring + slot, kF1SThreads);      // slot drained by the consumers
                int* sidx = reinterpret_cast<int*>(ring + slot * slot_bytes);
                float4* sd = reinterpret_cast<float4*>(ring + slot * slot_bytes + (size_t)kF1Batch * K * sizeof(int));
                const long long gq0 = q + (long long)bi * kF1Batch;                        // global query id of the batch
                const int nqb = min(kF1Batch, (int)(seg_end - gq0));
                // lane 3*i + t: query i of the batch, z-slab t
                const int qi = lane / kBqSlabLanes, tsl = lane - qi * kBqSlabLanes;
                const bool act = qi < nqb;
                float qx = 0.f, qy = 0.f, qz = 0.f;
                if (act) {
                    const float* p2 = a.new_xyz + (size_t)(gq0 + qi) * 3;
                    qx = __ldg(p2); qy = __ldg(p2 + 1); qz = __ldg(p2 + 2);
                }
                const bool qfin = fabsf(qx) <= 3.0e38f && fabsf(qy) <= 3.0e38f && fabsf(qz) <= 3.0e38f;
                const bool fast = g.use && !a.none;
                if (fast && act && qfin) bq_search_slab(s, g, a.thr, qx, qy, qz, tsl, bitmaps + (size_t)qi * 32 * wpl);
                __syncwarp();
                for (int i = 0; i < nqb; ++i) {
                    const float cx = __shfl_sync(0xffffffffu, qx, i * kBqSlabLanes), cy = __shfl_sync(0xffffffffu, qy, i * kBqSlabLanes);
                    const float cz = __shfl_sync(0xffffffffu, qz, i * kBqSlabLanes);
                    const bool fin_i = __shfl_sync(0xffffffffu, qfin ? 1 : 0, i * kBqSlabLanes) != 0;
                    int* row = sidx + i * K;
                    int cnt;
                    if (fast && fin_i) cnt = bq_extract_bitmap(bitmaps + (size_t)i * 32 * wpl, wpl, K, row, lane);
                    else cnt = bq_scan_warp(n, K, a.thr, a.none != 0, s, cx, cy, cz, row, lane);
                    __syncwarp();
                    if (a.pts_cnt != nullptr && lane == 0) a.pts_cnt[gq0 + i] = cnt;
                    // grouped_xyz - new_xyz (pointnet_util.py:46) of the query's K rows, + the source index for the U gather
                    int* gidx = a.idx + (size_t)(gq0 + i) * K;
                    for (int l = lane; l < K; l += 32) {
                        const int j = row[l];
                        gidx[l] = j;
                        const float px = __ldg(gx + 3 * j), py = __ldg(gx + 3 * j + 1), pz = __ldg(gx + 3 * j + 2);
                        sd[i * K + l] = make_float4(px - cx, py - cy, pz - cz, __int_as_float(j));
                    }
                }
                __threadfence_block();
                named_bar_arrive(1 + slot, kF1SThreads);                                    // FULL
            }
        } else {
            // =========================== CONSUMERS ===========================
            const int cw = warp - 1;
            // first-layer weights of this lane's channels, resident in registers (loaded after the grid build: the build
            // keeps the thread's points in registers)
            float2 wx[NPK], wy[NPK], wz[NPK], bs[NPK];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = 32 * i + 4 * sub;
                const float4 x4 = __ldg(reinterpret_cast<const float4*>(a.w1 + c));
                const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.w1 + C1 + c));
                const float4 z4 = __ldg(reinterpret_cast<const float4*>(a.w1 + 2 * C1 + c));
                const float4 b4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                wx[2 * i] = make_float2(x4.x, x4.y); wx[2 * i + 1] = make_float2(x4.z, x4.w);
                wy[2 * i] = make_float2(y4.x, y4.y); wy[2 * i + 1] = make_float2(y4.z, y4.w);
                wz[2 * i] = make_float2(z4.x, z4.y); wz[2 * i + 1] = make_float2(z4.z, z4.w);
                bs[2 * i] = make_float2(b4.x, b4.y); bs[2 * i + 1] = make_float2(b4.z, b4.w);
            }
            for (int bi = 0; bi < nb; ++bi, ++ring_pos) {
                const int slot = ring_pos % kF1Ring;
                const float4* sd = reinterpret_cast<const float4*>(ring + slot * slot_bytes + (size_t)kF1Batch * K * sizeof(int));
                const long long gq0 = q + (long long)bi * kF1Batch;
                const int nrows = min(kF1Batch, (int)(seg_end - gq0)) * K;
                float* outl = a.pre + (size_t)gq0 * K * C1 + 4 * sub;
                const float* ucloud = HAS_U ? a.uf + (size_t)cloud * n * C1 + 4 * sub : nullptr;
                named_bar_sync(1 + slot, kF1SThreads);                                      // FULL
                for (int r = 4 * cw + rsub; r < nrows; r += 4 * kF1SConsWarps) {
                    const float4 d = sd[r];
                    const float2 dx = make_float2(d.x, d.x), dy = make_float2(d.y, d.y), dz = make_float2(d.z, d.z);
                    float* orow = outl + (unsigned)r * (unsigned)C1;
                    const float* urow = HAS_U ? ucloud + (unsigned)__float_as_int(d.w) * (unsigned)C1 : nullptr;
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        float2 s0 = bs[2 * i], s1 = bs[2 * i + 1];
                        if (HAS_U) {
                            const float4 u = __ldg(reinterpret_cast<const float4*>(urow + 32 * i));
                            s0 = __fadd2_rn(s0, make_float2(u.x, u.y)); s1 = __fadd2_rn(s1, make_float2(u.z, u.w));
                        }
                        const float2 v0 = __ffma2_rn(dz, wz[2 * i], __ffma2_rn(dy, wy[2 * i], __ffma2_rn(dx, wx[2 * i], s0)));
                        const float2 v1 = __ffma2_rn(dz, wz[2 * i + 1], __ffma2_rn(dy, wy[2 * i + 1], __ffma2_rn(dx, wx[2 * i + 1], s1)));
                        __stcs(reinterpret_cast<float4*>(orow + 32 * i), make_float4(v0.x, v0.y, v1.x, v1.y));
                        if (STATS) {
                            ssum[2 * i] = __fadd2_rn(ssum[2 * i], v0); ssum[2 * i + 1] = __fadd2_rn(ssum[2 * i + 1], v1);
                            ssq[2 * i] = __ffma2_rn(v0, v0, ssq[2 * i]); ssq[2 * i + 1] = __ffma2_rn(v1, v1, ssq[2 * i + 1]);
                        }
                    }
                }
                if (ring_pos + kF1Ring < total_batches) named_bar_arrive(1 + kF1Ring + slot, kF1SThreads);   // EMPTY
            }
        }
        q = seg_end;
    }

    if (STATS) {
        __syncthreads();                                   // ring memory is free: reuse it for the per-warp partials
        float* sstat = reinterpret_cast<float*>(ring);     // kF1SConsWarps x 2 x C1
#pragma unroll
        for (int p = 0; p < NPK; ++p) {
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
                ssum[p].x += __shfl_xor_sync(0xffffffffu, ssum[p].x, o); ssum[p].y += __shfl_xor_sync(0xffffffffu, ssum[p].y, o);
                ssq[p].x += __shfl_xor_sync(0xffffffffu, ssq[p].x, o); ssq[p].y += __shfl_xor_sync(0xffffffffu, ssq[p].y, o);
            }
            if (warp > 0 && rsub == 0) {
                float* w = sstat + (size_t)(warp - 1) * 2 * C1;
                const int c = 32 * (p >> 1) + 4 * sub + 2 * (p & 1);
                *reinterpret_cast<float2*>(w + c) = ssum[p];
                *reinterpret_cast<float2*>(w + C1 + c) = ssq[p];
            }
        }
        __syncthreads();
        float* dst = a.partial + (size_t)blockIdx.x * 2 * C1;
        for (int e = tid; e < 2 * C1; e += kF1SThreads) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kF1SConsWarps; ++w) t += sstat[(size_t)w * 2 * C1 + e];      // fixed order
            dst[e] = t;
        }
        // last CTA to arrive adds the CTA partials in CTA order (fp64): deterministic whatever the finishing order
        __shared__ unsigned s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int e = tid; e < 2 * C1; e += kF1SThreads) {
                double t = 0.0;
                for (unsigned p = 0; p < gridDim.x; ++p) t += (double)__ldcg(a.partial + (size_t)p * 2 * C1 + e);
                a.stats[e] = (float)t;
            }
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
// 0 = streaming kernel where it applies, 1 = round-1 kernel; PSA_F1_VARIANT in the environment (A/B runs of tools/ only)
static int f1_variant() {
    static const int v = [] { const char* e = getenv("PSA_F1_VARIANT"); return e ? atoi(e) : 0; }();
    return v;
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

// streaming kernel: grid, shared memory, applicability
static bool f1s_plan(int b, int n, int m, int nsample, bool want_grid, int* ctas, size_t* smem) {
    if (n > (want_grid ? kBqGridMaxN : 8192) || nsample > 128) return false;
    *smem = f1s_smem_bytes(n, nsample, want_grid);
    if (*smem > 200 * 1024) return false;
    const long long T = (long long)b * m;
    const long long batches = (T + kF1Batch - 1) / kF1Batch;
    const int per_sm = (*smem <= 110 * 1024) ? 2 : 1;
    *ctas = (int)(batches < (long long)per_sm * kNumSMs ? batches : (long long)per_sm * kNumSMs);
    return true;
}
static bool f1_want_grid(int n, int m) { return bq_grid_fits(n) && n >= 256 && m >= 32; }

extern "C" size_t psa_sa_conv1_prebn_workspace_bytes(int b, int n, int m, int c, int C1, int want_stats) {
    size_t bytes = 0;
    if (c > 0) bytes += ((size_t)b * n * C1 * sizeof(float) + 255) & ~(size_t)255;
    if (want_stats) {
        int q; dim3 g;
        f1_grid(b, m, &q, &g);
        size_t parts = (size_t)g.x * g.y;
        if (parts < 2 * (size_t)kNumSMs) parts = 2 * (size_t)kNumSMs;
        bytes += parts * 2 * C1 * sizeof(float) + 256;          // CTA partials + the completion ticket
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
    a.want_grid = f1_want_grid(n, m) ? 1 : 0;
    a.xyz = xyz; a.new_xyz = new_xyz; a.w1 = w1; a.bias = bias; a.pre = pre; a.idx = idx; a.pts_cnt = pts_cnt;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    a.uf = nullptr;
    if (c > 0) {
        int rc = launch_dense_raw((long long)b * n, c, C1, points, w1 + (size_t)3 * C1, reinterpret_cast<float*>(ws), st);
        if (rc != PSA_OK) return rc;
        a.uf = reinterpret_cast<float*>(ws);
        ws += ((size_t)b * n * C1 * sizeof(float) + 255) & ~(size_t)255;
    }
    {   // ---- streaming kernel (persistent CTAs, producer/consumer warps) whenever the cloud fits its shared-memory plan ----
        int ctas = 0;
        size_t ssm = 0;
        if (f1_variant() != 1 && f1s_plan(b, n, m, nsample, a.want_grid != 0, &ctas, &ssm)) {
            F1SArgs s;
            s.b = b; s.n = n; s.m = m; s.nsample = nsample; s.C1 = C1; s.radius = radius; s.thr = a.thr; s.none = a.none;
            s.want_grid = a.want_grid; s.xyz = xyz; s.new_xyz = new_xyz; s.uf = a.uf; s.w1 = w1; s.bias = bias; s.pre = pre;
            s.idx = idx; s.pts_cnt = pts_cnt; s.partial = nullptr; s.stats = stats; s.ticket = nullptr;
            if (stats) {
                s.ticket = reinterpret_cast<unsigned*>(ws);
                s.partial = reinterpret_cast<float*>(ws + 256);
                PSA_CUDA(cudaMemsetAsync(s.ticket, 0, sizeof(unsigned), st));
            }
#define PSA_F1S_LAUNCH(NV_, ST_, U_, PPT_)                                                                                    \
    do {                                                                                                                     \
        PSA_CUDA(cudaFuncSetAttribute(sa_conv1_stream_kernel<NV_, ST_, U_, PPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm)); \
        sa_conv1_stream_kernel<NV_, ST_, U_, PPT_><<<ctas, kF1SThreads, ssm, st>>>(s);                                       \
    } while (0)
#define PSA_F1S_U(NV_, ST_, PPT_) do { if (s.uf) PSA_F1S_LAUNCH(NV_, ST_, true, PPT_); else PSA_F1S_LAUNCH(NV_, ST_, false, PPT_); } while (0)
#define PSA_F1S_ST(NV_, PPT_) do { if (stats) PSA_F1S_U(NV_, true, PPT_); else PSA_F1S_U(NV_, false, PPT_); } while (0)
            if (n <= 8 * kBqThreads) { if (C1 == 64) PSA_F1S_ST(2, 8); else PSA_F1S_ST(4, 8); }
            else { if (C1 == 64) PSA_F1S_ST(2, 16); else PSA_F1S_ST(4, 16); }
#undef PSA_F1S_ST
#undef PSA_F1S_U
#undef PSA_F1S_LAUNCH
            return check_launch("sa_conv1_stream_kernel");
        }
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

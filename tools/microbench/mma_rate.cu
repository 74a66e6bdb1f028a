// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M = 128) as a function of N, kind, A source and issue style.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../scanobjectnn_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace psa::tc;

// warp-converged issue: every lane executes the instruction stream, one elected lane issues the MMA
template <int TF32>
__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (TF32)
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}

// style 0: `if (tid == 0)` single-thread issue (what the kernels did); style 1: warp-converged + elect, fully unrolled by 8
template <int TF32, int STYLE>
__global__ void __launch_bounds__(128, 1) k(int N, int count, int nd, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t idesc = make_idesc(TF32 ? kFmtTF32 : kFmtBF16, 128, N);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(base));
    if (STYLE == 0) {
        if (tid == 0) {
            long long t0 = clock64();
            for (int i = 0; i < count; ++i) {
                if (TF32) mma_tf32_ts(tb, tb + 256 + (i & 7) * 8, bdesc0 + (uint64_t)((i & 3) * 2), idesc, i > 0);
                else mma_bf16_ts(tb, tb + 256 + (i & 7) * 8, bdesc0 + (uint64_t)((i & 3) * 2), idesc, i > 0);
            }
            long long t1 = clock64();
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
    } else {
        if (warp == 0) {
            long long t0 = clock64();
            for (int i0 = 0; i0 < count; i0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t dd = tb + (uint32_t)((nd == 2) ? (u & 1) * 128 : 0);
                    mma_ts_elect<TF32>(dd, tb + 256 + u * 8, bdesc0 + (uint64_t)((u & 3) * 2), idesc, (i0 > 0 || u >= nd) ? 1u : 0u);
                }
            }
            long long t1 = clock64();
            commit_elect(&bar);
            mbar_wait(&bar, 0);
            long long t2 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int TF32, int STYLE>
void run(const char* name, long long* d) {
    cudaFuncSetAttribute(k<TF32, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int N : {64, 128, 256})
        for (int nd : {1, 2}) {
            if (nd * N > 256 || (STYLE == 0 && nd > 1)) continue;
            const int count = 512;
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) { k<TF32, STYLE><<<1, 128, 100 * 1024>>>(N, count, nd, d); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); }
            cudaError_t e = cudaGetLastError();
            printf("%-28s N=%3d nd=%d: issue %6.1f complete %6.1f cyc/mma (tensor-bound: %5.1f) %s\n", name, N, nd, (double)h[0] / count, (double)h[1] / count,
                   2.0 * 128 * N * 16 / 8192.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    run<0, 0>("TS bf16 K16 single-thread", d);
    run<1, 0>("TS tf32 K8  single-thread", d);
    run<0, 1>("TS bf16 K16 warp+elect", d);
    run<1, 1>("TS tf32 K8  warp+elect", d);
    return 0;
}

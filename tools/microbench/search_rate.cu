// search_rate.cu -- cycles per (query, warp) of the F1 ball search inner loop: 16 register-resident points per lane against one query,
// packed f32x2 (FADD2/FMUL2/FFMA2 + sign funnel shifts) vs scalar FADD/FMUL/FFMA, for 1 / 2 / 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/search_rate search_rate.cu && /tmp/search_rate
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void unpk(u64 v, unsigned& lo, unsigned& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }

template <int MODE>
__global__ void __launch_bounds__(512) k(const float* __restrict__ p, const float* __restrict__ q, unsigned* __restrict__ out, float thr, int nq, long long* cyc) {
    const int t = threadIdx.x;
    float x[16], y[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float* b = p + (t * 16 + i) * 3; x[i] = b[0]; y[i] = b[1]; z[i] = b[2]; }
    u64 px[8], py[8], pz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { px[j] = pk2(x[2 * j], x[2 * j + 1]); py[j] = pk2(y[2 * j], y[2 * j + 1]); pz[j] = pk2(z[2 * j], z[2 * j + 1]); }
    const u64 thr2 = pk2(thr, thr);
    unsigned sum = 0;
    __syncthreads();
    const long long c0 = clock64();
    for (int qi = 0; qi < nq; ++qi) {
        const float qx = q[3 * qi], qy = q[3 * qi + 1], qz = q[3 * qi + 2];
        unsigned a0 = 0, a1 = 0;
        if (MODE == 0) {
            const u64 nx = pk2(-qx, -qx), ny = pk2(-qy, -qy), nz = pk2(-qz, -qz);
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                const u64 dx = add2(px[j], nx), dy = add2(py[j], ny), dz = add2(pz[j], nz);
                u64 tt = mul2(dy, dy); tt = fma2(dx, dx, tt); tt = fma2(dz, dz, tt);
                unsigned lo, hi; unpk(sub2(thr2, tt), lo, hi);
                if (j >= 4) { a1 = __funnelshift_l(hi, a1, 1); a1 = __funnelshift_l(lo, a1, 1); }
                else { a0 = __funnelshift_l(hi, a0, 1); a0 = __funnelshift_l(lo, a0, 1); }
            }
        } else {
#pragma unroll
            for (int i = 15; i >= 0; --i) {
                const float dx = x[i] - qx, dy = y[i] - qy, dz = z[i] - qz;
                float tt = dy * dy; tt = fmaf(dx, dx, tt); tt = fmaf(dz, dz, tt);
                const unsigned s = __float_as_uint(thr - tt);
                if (i >= 8) a1 = __funnelshift_l(s, a1, 1); else a0 = __funnelshift_l(s, a0, 1);
            }
        }
        sum += ~((a1 << 8) | a0) & 0xffffu;
    }
    const long long c1 = clock64();
    out[blockIdx.x * blockDim.x + t] = sum;
    if (t == 0) cyc[blockIdx.x] = c1 - c0;
}

int main() {
    float *p, *q; unsigned* out; long long* cyc;
    cudaMalloc(&p, 512 * 16 * 3 * 4); cudaMalloc(&q, 4096 * 3 * 4); cudaMalloc(&out, 148 * 2 * 512 * 4); cudaMalloc(&cyc, 148 * 2 * 8);
    cudaMemset(p, 0, 512 * 16 * 3 * 4); cudaMemset(q, 0, 4096 * 3 * 4);
    const int nq = 2048;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads = 128; threads <= 512; threads *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(p, q, out, 0.04f, nq, cyc); else k<1><<<148, threads>>>(p, q, out, 0.04f, nq, cyc);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i]; m /= 148;
            printf("{\"mode\": \"%s\", \"warps_per_scheduler\": %d, \"cycles_per_query_per_warp\": %.1f, \"cycles_per_query_per_scheduler\": %.1f}\n",
                   mode == 0 ? "packed f32x2" : "scalar", threads / 128, m / nq, m / nq / (threads / 128));
        }
    return 0;
}

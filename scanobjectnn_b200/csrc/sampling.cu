// sampling.cu -- farthest point sampling + gather_point (+grad) for sm_100a.
//
// Replaces pointnet2/tf_ops/sampling/tf_sampling_g.cu:105-192 of the reference.  Results are index-exact
// with the reference kernel, including its tie-break: thread t of the reference's 512-thread block scans
// k = t, t+512, ... with strict '>', then a tree that keeps the lower slot on ties, i.e. the winner among
// equal maxima is min over (k mod 512, k).  This kernel keeps the same thread<->point mapping (512 threads,
// point k owned by thread k mod 512) so "first maximum in (lane, warp) order" IS that key -- but the running
// min-distances and coordinates live in registers (no global `temp` round trip), the block arg-max is one
// REDUX + ballot per warp and one smem hop (1 __syncthreads per round instead of the reference's 10), and
// the gather of the sampled coordinates is fused.
#include <limits.h>

#include "common.cuh"

namespace psa {

constexpr int kFpsThreads = 512;   // == the reference's BlockSize: defines the tie-break key
constexpr int kFpsWarps = kFpsThreads / 32;

template <int PPT>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_kernel(int n, int m, const float* __restrict__ xyz, int* __restrict__ idx_out, float* __restrict__ new_xyz) {
    extern __shared__ float smem_f[];
    float* sxyz = smem_f;                                  // n*3 floats, flat copy of the cloud
    int* s_sel = reinterpret_cast<int*>(smem_f + (size_t)n * 3);   // m selected indices
    __shared__ int s_wval[2][kFpsWarps];
    __shared__ int s_widx[2][kFpsWarps];

    const int cloud = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const float* p = xyz + (size_t)cloud * n * 3;

    for (int i = t; i < n * 3; i += kFpsThreads) sxyz[i] = p[i];
    __syncthreads();

    float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int k = t + i * kFpsThreads;
        if (k < n) {
            px[i] = sxyz[k * 3 + 0]; py[i] = sxyz[k * 3 + 1]; pz[i] = sxyz[k * 3 + 2];
            td[i] = 1e38f;                                 // tf_sampling_g.cu:118
        } else {
            px[i] = py[i] = pz[i] = 0.f;
            td[i] = -1.f;                                  // never beats best=-1 (strict >), fminf keeps it
        }
    }

    int old = 0;                                           // seed index 0 (tf_sampling_g.cu:114)
    if (t == 0) s_sel[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = sxyz[old * 3 + 0], y1 = sxyz[old * 3 + 1], z1 = sxyz[old * 3 + 2];
        float best = -1.f;
        int besti = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float d = dist2_ref_gpu(px[i] - x1, py[i] - y1, pz[i] - z1);
            float d2 = fminf(d, td[i]);                    // NaN d leaves td unchanged, as CUDA min() does
            td[i] = d2;
            if (d2 > best) { best = d2; besti = t + i * kFpsThreads; }
        }
        // best is -1 or a non-negative float (incl. +inf): signed-int order == float order
        const int key = __float_as_int(best);
        const int wmax = __reduce_max_sync(0xffffffffu, key);
        const unsigned bal = __ballot_sync(0xffffffffu, key == wmax);
        const int src = __ffs(bal) - 1;                    // lowest lane == lowest slot wins ties
        const int widx = __shfl_sync(0xffffffffu, besti, src);
        const int par = j & 1;
        if (lane == 0) { s_wval[par][warp] = wmax; s_widx[par][warp] = widx; }
        __syncthreads();
        const int v = (lane < kFpsWarps) ? s_wval[par][lane] : INT_MIN;
        const int gmax = __reduce_max_sync(0xffffffffu, v);
        const unsigned bal2 = __ballot_sync(0xffffffffu, v == gmax);
        old = s_widx[par][__ffs(bal2) - 1];                // lowest warp == lowest slot range wins ties
        if (t == 0) s_sel[j] = old;
    }
    __syncthreads();
    int* io = idx_out + (size_t)cloud * m;
    for (int j = t; j < m; j += kFpsThreads) io[j] = s_sel[j];
    if (new_xyz != nullptr) {
        float* o = new_xyz + (size_t)cloud * m * 3;
        for (int e = t; e < m * 3; e += kFpsThreads) {
            int j = e / 3, c = e - j * 3;
            o[e] = sxyz[s_sel[j] * 3 + c];
        }
    }
}

__global__ void gather_point_kernel(int n, int m, long long total, const float* __restrict__ inp,
                                    const int* __restrict__ idx, float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long row = e / 3;
        int c = (int)(e - row * 3);
        long long bi = row / m;
        int a = idx[row];
        out[e] = inp[(bi * n + a) * 3 + c];
    }
}

__global__ void gather_point_grad_kernel(int n, int m, long long total, const float* __restrict__ out_g,
                                         const int* __restrict__ idx, float* __restrict__ inp_g) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long row = e / 3;
        int c = (int)(e - row * 3);
        long long bi = row / m;
        int a = idx[row];
        atomicAdd(&inp_g[(bi * n + a) * 3 + c], out_g[e]);
    }
}

template <int PPT>
static int launch_fps(int b, int n, int m, const float* xyz, int* idx, float* new_xyz, cudaStream_t st) {
    size_t smem = (size_t)n * 3 * sizeof(float) + (size_t)m * sizeof(int);
    PSA_SUPPORTED(smem <= 200 * 1024, "farthest_point_sample: n=%d, m=%d needs %zu B of shared memory", n, m, smem);
    PSA_CUDA(cudaFuncSetAttribute(fps_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_kernel<PPT><<<b, kFpsThreads, smem, st>>>(n, m, xyz, idx, new_xyz);
    return check_launch("fps_kernel");
}

static inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    long long cap = (long long)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace psa

using namespace psa;

extern "C" int psa_farthest_point_sample(int b, int n, int m, const float* xyz, int* idx, float* new_xyz,
                                         psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "FarthestPointSample: negative dimension (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || m == 0) return PSA_OK;
    PSA_REQUIRE(n >= 1, "FarthestPointSample expects at least one input point (n=%d)", n);
    PSA_REQUIRE(xyz != nullptr && idx != nullptr, "FarthestPointSample: null buffer");
    cudaStream_t st = as_stream(stream);
    const int ppt = (n + kFpsThreads - 1) / kFpsThreads;
    if (ppt <= 1) return launch_fps<1>(b, n, m, xyz, idx, new_xyz, st);
    if (ppt <= 2) return launch_fps<2>(b, n, m, xyz, idx, new_xyz, st);
    if (ppt <= 4) return launch_fps<4>(b, n, m, xyz, idx, new_xyz, st);
    if (ppt <= 8) return launch_fps<8>(b, n, m, xyz, idx, new_xyz, st);
    if (ppt <= 16) return launch_fps<16>(b, n, m, xyz, idx, new_xyz, st);
    PSA_SUPPORTED(false, "farthest_point_sample: n=%d exceeds the register-resident limit of %d points", n,
                  16 * kFpsThreads);
}

extern "C" int psa_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out,
                                psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "GatherPoint: negative dimension");
    long long total = (long long)b * m * 3;
    if (total == 0) return PSA_OK;
    PSA_REQUIRE(inp && idx && out, "GatherPoint: null buffer");
    gather_point_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(n, m, total, inp, idx, out);
    return check_launch("gather_point_kernel");
}

extern "C" int psa_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g,
                                     psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "GatherPointGrad: negative dimension");
    if ((long long)b * n == 0) return PSA_OK;
    PSA_REQUIRE(inp_g != nullptr, "GatherPointGrad: null buffer");
    PSA_CUDA(cudaMemsetAsync(inp_g, 0, sizeof(float) * (size_t)b * n * 3, as_stream(stream)));
    long long total = (long long)b * m * 3;
    if (total == 0) return PSA_OK;
    PSA_REQUIRE(out_g && idx, "GatherPointGrad: null buffer");
    gather_point_grad_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(n, m, total, out_g, idx, inp_g);
    return check_launch("gather_point_grad_kernel");
}

// augment.cu -- the reference's per-epoch / per-batch input pipeline as ONE kernel.
//
// What the reference does on the host with numpy, once per epoch and once per batch, before feed_dict + H2D:
//   data_utils.py:162-168  center_data      pc -= mean(pc, axis=0)                       (over ALL source points)
//   data_utils.py:133-143  normalize_data   pc /= max_i sqrt(x_i^2 + y_i^2 + z_i^2)
//   data_utils.py:171-186  get_current_data_h5   the same random point subset (idx_pts[:num_points]) for every cloud
//   provider.py:34-52      rotate_point_cloud    p . [[c,0,s],[0,1,0],[-s,0,c]]  (float64 product, stored as float32)
//   provider.py:215-227    random_scale_point_cloud, shift_point_cloud (per cloud)
//   provider.py:189-200    jitter_point_cloud     p + clip(sigma * randn, -clip, clip)   (float64 sum, fed as float32)
//   provider.py:229-236    random_point_dropout   dropped points := point 0
// Order inside the kernel: dropout substitution -> rotate -> scale -> shift -> jitter.  rotate -> jitter is the composition the
// training scripts run (pointnet2/train.py:246-252); the order of the optional scale / shift / dropout relative to the jitter is
// this kernel's own (the reference composes them only in commented-out code, dgcnn/train.py:274-278, jitter first).
// Here: one CTA per cloud, pass 1 (optional) reduces the centroid and the max norm over the source cloud, pass 2
// gathers the subset and applies the chain in the reference's order and precisions, writing the (b,n,3) batch the
// first FPS reads.  Random numbers are INPUTS (permutation, angles as cos/sin in double, scales, shifts, standard
// normal noise, dropout mask): the caller draws them (torch on device, or numpy for parity tests); the arithmetic is here.
#include <float.h>

#include "common.cuh"

namespace psa {

constexpr int kAugThreads = 256;

struct AugArgs {
    int n_src, n;
    const float* src;        // (b, n_src, 3)
    const int* perm;         // (n) shared by all clouds, or null = first n points
    const double* cs;        // (b, 2) cos, sin of the rotation about the up (y) axis, or null
    const float* scale;      // (b) or null
    const float* shift;      // (b, 3) or null
    const float* noise;      // (b, n, 3) standard normal, or null
    const unsigned char* drop;   // (b, n) 1 = dropped (replaced by the cloud's first output point before scaling), or null
    double sigma, clip;      // Python floats in the reference (float64)
    int center, normalize;
    float* out;              // (b, n, 3)
};

__global__ void __launch_bounds__(kAugThreads)
augment_kernel(const AugArgs a) {
    __shared__ float s_red[4][kAugThreads / 32];
    __shared__ float s_stat[4];          // centroid xyz, 1 / max norm is NOT stored: the division is by d itself
    const int bi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* src = a.src + (size_t)bi * a.n_src * 3;
    float cx = 0.f, cy = 0.f, cz = 0.f, d = 1.f;
    if (a.center) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int i = tid; i < a.n_src; i += kAugThreads) { sx += __ldg(src + 3 * i); sy += __ldg(src + 3 * i + 1); sz += __ldg(src + 3 * i + 2); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
        }
        if (lane == 0) { s_red[0][warp] = sx; s_red[1][warp] = sy; s_red[2][warp] = sz; }
        __syncthreads();
        if (tid == 0) {
            float tx = 0.f, ty = 0.f, tz = 0.f;
            for (int w = 0; w < kAugThreads / 32; ++w) { tx += s_red[0][w]; ty += s_red[1][w]; tz += s_red[2][w]; }
            s_stat[0] = tx / (float)a.n_src; s_stat[1] = ty / (float)a.n_src; s_stat[2] = tz / (float)a.n_src;
        }
        __syncthreads();
        cx = s_stat[0]; cy = s_stat[1]; cz = s_stat[2];
    }
    if (a.normalize) {
        float mx = 0.f;
        for (int i = tid; i < a.n_src; i += kAugThreads) {
            const float x = __fsub_rn(__ldg(src + 3 * i), cx), y = __fsub_rn(__ldg(src + 3 * i + 1), cy), z = __fsub_rn(__ldg(src + 3 * i + 2), cz);
            mx = fmaxf(mx, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[3][warp] = mx;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < kAugThreads / 32; ++w) t = fmaxf(t, s_red[3][w]);
            s_stat[3] = t;
        }
        __syncthreads();
        d = s_stat[3];
    }
    double c = 1.0, s = 0.0;
    if (a.cs) { c = a.cs[2 * bi]; s = a.cs[2 * bi + 1]; }
    const float sc = a.scale ? __ldg(a.scale + bi) : 1.f;
    const float hx = a.shift ? __ldg(a.shift + 3 * bi) : 0.f, hy = a.shift ? __ldg(a.shift + 3 * bi + 1) : 0.f, hz = a.shift ? __ldg(a.shift + 3 * bi + 2) : 0.f;
    float* out = a.out + (size_t)bi * a.n * 3;
    for (int i = tid; i < a.n; i += kAugThreads) {
        int j = i;
        if (a.drop && a.drop[(size_t)bi * a.n + i]) j = 0;             // random_point_dropout: := the first point of the batch cloud
        const int sj = a.perm ? __ldg(a.perm + j) : j;
        float x = __fsub_rn(__ldg(src + 3 * sj), cx), y = __fsub_rn(__ldg(src + 3 * sj + 1), cy), z = __fsub_rn(__ldg(src + 3 * sj + 2), cz);
        if (a.normalize) { x = __fdiv_rn(x, d); y = __fdiv_rn(y, d); z = __fdiv_rn(z, d); }
        if (a.cs) {
            // np.dot(float32 (n,3), float64 (3,3)) -> float64, stored into a float32 array
            const double xd = x, yd = y, zd = z;
            const double rx = (xd * c + yd * 0.0) + zd * (-s);
            const double ry = (xd * 0.0 + yd * 1.0) + zd * 0.0;
            const double rz = (xd * s + yd * 0.0) + zd * c;
            x = (float)rx; y = (float)ry; z = (float)rz;
        }
        if (a.scale) { x = __fmul_rn(x, sc); y = __fmul_rn(y, sc); z = __fmul_rn(z, sc); }
        if (a.shift) { x = __fadd_rn(x, hx); y = __fadd_rn(y, hy); z = __fadd_rn(z, hz); }
        if (a.noise) {
            // float64 jitter added to the float32 data in float64, then fed as float32 (provider.py:197-199)
            const float* nz = a.noise + ((size_t)bi * a.n + i) * 3;
            const double sg = a.sigma, cl = a.clip;
            const double jx = fmin(fmax(sg * (double)__ldg(nz), -cl), cl), jy = fmin(fmax(sg * (double)__ldg(nz + 1), -cl), cl),
                         jz = fmin(fmax(sg * (double)__ldg(nz + 2), -cl), cl);
            x = (float)(jx + (double)x); y = (float)(jy + (double)y); z = (float)(jz + (double)z);
        }
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
}

}  // namespace psa

using namespace psa;

extern "C" int psa_augment_batch(int b, int n_src, int n, const float* src, const int* perm, const double* cos_sin,
                                 const float* scale, const float* shift, const float* noise, double sigma, double clip,
                                 const unsigned char* drop, int center, int normalize, float* out, psa_stream_t stream) {
    PSA_REQUIRE(b >= 0 && n_src >= 1 && n >= 0, "augment_batch: bad dims b=%d n_src=%d n=%d", b, n_src, n);
    PSA_REQUIRE(perm != nullptr || n <= n_src, "augment_batch: n=%d exceeds the source cloud (%d points) and no index list is given", n, n_src);
    PSA_REQUIRE(noise == nullptr || clip > 0.0, "augment_batch: clip must be positive (provider.py:196)");
    if (b == 0 || n == 0) return PSA_OK;
    PSA_REQUIRE(src && out, "augment_batch: null buffer");
    AugArgs a;
    a.n_src = n_src; a.n = n; a.src = src; a.perm = perm; a.cs = cos_sin; a.scale = scale; a.shift = shift; a.noise = noise;
    a.drop = drop; a.sigma = sigma; a.clip = clip; a.center = center; a.normalize = normalize; a.out = out;
    augment_kernel<<<b, kAugThreads, 0, as_stream(stream)>>>(a);
    return check_launch("augment_kernel");
}

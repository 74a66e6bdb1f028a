// extern "C" doorway onto the reference's CPU three_nn / three_interpolate op bodies
// (pointnet2/tf_ops/3d_interpolation/tf_interpolate.cpp:57-153).  That file also includes TensorFlow
// headers and op classes, so the Makefile slices the TF-free lines 57-153 into the git-ignored
// oracle/_ref/tf_interpolate_slice.inc at build time (generated, never committed) and this shim
// includes the slice.  TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstring>
#include "tf_interpolate_slice.inc"
extern "C" {
void refcpu_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx) {
    threenn_cpu(b, n, m, xyz1, xyz2, dist, idx);
}
void refcpu_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx,
                              const float* weight, float* out) {
    threeinterpolate_cpu(b, m, c, n, points, idx, weight, out);
}
// caller zeroes grad_points (tf_interpolate.cpp:258)
void refcpu_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx,
                                   const float* weight, float* grad_points) {
    threeinterpolate_grad_cpu(b, n, c, m, grad_out, idx, weight, grad_points);
}
}

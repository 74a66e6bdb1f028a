// extern "C" doorway onto the reference's CPU ball-query / group harness functions
// (pointnet2/tf_ops/grouping/test/query_ball_point.cpp:19-84), compiled where the file lies (-I$(REF)/...).
// TEST INFRASTRUCTURE ONLY.  main()/randomf() of the harness are renamed out of the way by the Makefile.
#include "query_ball_point.cpp"
extern "C" {
void refcpu_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1,
                             const float* xyz2, int* idx) {
    query_ball_point_cpu(b, n, m, radius, nsample, xyz1, xyz2, idx);
}
void refcpu_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out) {
    group_point_cpu(b, n, c, m, nsample, points, idx, out);
}
// caller zeroes grad_points
void refcpu_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                             float* grad_points) {
    group_point_grad_cpu(b, n, c, m, nsample, grad_out, idx, grad_points);
}
}

// ref_gpu_shim.cu -- extern "C" doorway onto the REFERENCE's own CUDA launchers.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  No reference source is copied: the Makefile
// compiles pointnet2/tf_ops/{sampling/tf_sampling_g.cu,grouping/tf_grouping_g.cu} where they lie under
// $(REF) into objects under oracle/_ref/ and links them with this file into oracle/_ref/libref_tfops.so.
// The launchers are declared exactly as the reference's op glue declares them
// (tf_sampling.cpp:65,94,125,150 ; tf_grouping.cpp:66,108,142,173).
#include <cuda_runtime.h>

void farthestpointsamplingLauncher(int b, int n, int m, const float* inp, float* temp, int* out);
void gatherpointLauncher(int b, int n, int m, const float* inp, const int* idx, float* out);
void scatteraddpointLauncher(int b, int n, int m, const float* out_g, const int* idx, float* inp_g);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float* xyz1,
                            const float* xyz2, int* idx, int* pts_cnt);
void selectionSortLauncher(int b, int n, int m, int k, const float* dist, int* outi, float* out);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float* points, const int* idx,
                        float* out);
void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                            float* grad_points);

static int finish(int sync) {
    if (sync) return (int)cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}

extern "C" {
// temp must hold 32*n floats (tf_sampling.cpp:115)
int ref_fps(int b, int n, int m, const float* inp, float* temp, int* out, int sync) {
    farthestpointsamplingLauncher(b, n, m, inp, temp, out);
    return finish(sync);
}
int ref_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out, int sync) {
    gatherpointLauncher(b, n, m, inp, idx, out);
    return finish(sync);
}
// caller zeroes inp_g first, as tf_sampling.cpp:174 does
int ref_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g, int sync) {
    scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
    return finish(sync);
}
int ref_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                         int* idx, int* pts_cnt, int sync) {
    queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
    return finish(sync);
}
int ref_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out, int sync) {
    selectionSortLauncher(b, n, m, k, dist, outi, out);
    return finish(sync);
}
int ref_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                    int sync) {
    groupPointLauncher(b, n, c, m, nsample, points, idx, out);
    return finish(sync);
}
// caller zeroes grad_points first, as tf_grouping.cpp:204 does
int ref_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                         float* grad_points, int sync) {
    groupPointGradLauncher(b, n, c, m, nsample, grad_out, idx, grad_points);
    return finish(sync);
}
}

// extern "C" doorway onto the reference's CPU selection sort harness
// (pointnet2/tf_ops/grouping/test/selection_sort.cpp:19-64), compiled where the file lies.
// TEST INFRASTRUCTURE ONLY.  The harness prints inside its loops (:36-38,:48); printf is muted here.
#include <cstdio>
#define printf(...) ((void)0)
#include "selection_sort.cpp"
#undef printf
extern "C" {
void refcpu_selection_sort(int b, int n, int m, int k, const float* dist, int* idx, float* val) {
    selection_sort_cpu(b, n, m, k, dist, idx, val);
}
}

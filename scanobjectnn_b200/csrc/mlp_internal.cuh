// mlp_internal.cuh -- pieces of mlp.cu shared with tc_mlp.cu
#pragma once
#include "common.cuh"

namespace psa {

struct DenseArgs {
    long long rows;
    int K, N;
    int pool_k;          // 1 = none
    int relu;
    const float* x;      // (rows, K)
    const float* W;      // (K, N)
    const float* scale;  // (N) or null
    const float* shift;  // (N)
    float* out;          // (rows, N) or (rows/pool_k, N)
};


int launch_dense(const DenseArgs& d, cudaStream_t st);
size_t fc_small_workspace_bytes(int K, int N);
int launch_fc_small(const DenseArgs& d, float* partial, cudaStream_t st);   // rows <= 32
int sa_module_simt(int b, int n, int m, int c, int nsample, const float* xyz, const float* new_xyz, const float* points,
                   const int* idx, const psa_mlp* mlp, float* out, cudaStream_t st);
int validate_mlp_public(const psa_mlp* mlp, const char* who);
int edgeconv_simt(int b, int n, int c, int k, const float* x, const int* nn_idx, const psa_mlp* mlp, float* out, cudaStream_t st);
// `run_if` non-null: no-op unless *run_if != 0 (device-side condition, see tc_mlp.cu)
int launch_fill_ord_neg_inf(long long total, float* out, cudaStream_t st, const unsigned int* run_if = nullptr);   // out := order-preserving int code of -inf
int launch_decode_ord(long long total, float* out, cudaStream_t st, const unsigned int* run_if = nullptr);         // int codes -> floats, in place

}  // namespace psa

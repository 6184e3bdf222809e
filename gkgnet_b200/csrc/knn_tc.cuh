// tcgen05 / TMEM kNN kernel: host-side interface (see knn_tc.cu).
#pragma once
#include "common.cuh"

namespace gkg {

bool knn_tc_supported(int N, int M, int D, int k, int dilation);
bool knn_tc_preferred(int N, int M, int D, int k, int dilation);
size_t knn_tc_workspace_bytes(int P, int N, int M, int D, int k, int dilation, bool self_keys);
int launch_knn_tc_prepare(const KnnWorkspace& w, void* extra_ws, const void* x, int64_t x_sb, int64_t x_sn,
                          const void* y, int64_t y_sb, int64_t y_sn, int dtype, int P, int G, int N, int M,
                          int D, int k, int dilation, bool self_keys, cudaStream_t stream);
// separable form of the position bias: relpos[n, m] = a[n % grid_w][m % kw] + b[n / grid_w][m / kw]
struct SepBias {
  const float* a;
  const float* b;
  int grid_w, kw;
};
int launch_knn_tc(const KnnWorkspace& w, void* extra_ws, const float* relpos, const SepBias& sep,
                  int32_t* idx_out, int P, int N, int M, int D, int k, int dilation, bool self_keys,
                  cudaStream_t stream);

}  // namespace gkg

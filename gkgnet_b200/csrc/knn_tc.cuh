// tcgen05 / TMEM kNN kernel: host-side interface (see knn_tc.cu).
#pragma once
#include "common.cuh"

namespace gkg {

bool knn_tc_supported(int N, int M, int D, int k, int dilation);
size_t knn_tc_workspace_bytes(int P, int N, int M, int D, int k, int dilation, bool self_keys);
int launch_knn_tc_prepare(const KnnWorkspace& w, void* extra_ws, int P, int N, int M, int D, int k,
                          int dilation, bool self_keys, cudaStream_t stream);
int launch_knn_tc(const KnnWorkspace& w, void* extra_ws, const float* relpos, int32_t* idx_out, int P,
                  int N, int M, int D, int k, int dilation, bool self_keys, cudaStream_t stream);

}  // namespace gkg

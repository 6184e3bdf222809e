// placeholder until the tcgen05 kernel lands
#include "knn_tc.cuh"
namespace gkg {
bool knn_tc_supported(int, int, int, int, int) { return false; }
size_t knn_tc_workspace_bytes(int, int, int, int, int, int, bool) { return 0; }
int launch_knn_tc_prepare(const KnnWorkspace&, void*, int, int, int, int, int, int, bool, cudaStream_t) {
  set_error("knn_tc: not built");
  return GKG_EINVAL;
}
int launch_knn_tc(const KnnWorkspace&, void*, const float*, int32_t*, int, int, int, int, int, int, bool,
                  cudaStream_t) {
  set_error("knn_tc: not built");
  return GKG_EINVAL;
}
}  // namespace gkg

// tcgen05 / TMEM kNN kernel: host-side interface (see knn_tc.cu).
#pragma once
#include <atomic>

#include "common.cuh"

namespace gkg {

bool knn_tc_supported(int N, int M, int D, int k, int dilation, int dtype);
bool knn_tc_preferred(int N, int M, int D, int k, int dilation, int dtype);
size_t knn_tc_workspace_bytes(int P, int N, int M, int D, int k, int dilation, int dtype);
int launch_knn_tc_prepare(const KnnWorkspace& w, void* extra_ws, const void* x, int64_t x_sb, int64_t x_sn,
                          const void* y, int64_t y_sb, int64_t y_sn, int dtype, int P, int G, int N, int M,
                          int D, int k, int dilation, bool self_keys, cudaStream_t stream);
// separable form of the position bias: relpos[n, m] = a[n % grid_w][m % kw] + b[n / grid_w][m / kw]
struct SepBias {
  const float* a;
  const float* b;
  int grid_w, kw;
};
// per-call test hooks (gkg_knn_select_debug); the product entry points pass none
struct KnnDebug {
  int flags;                 // 1: re-rank every row exactly, 2: MMA only (no ranking), 3: every row to the fix-up kernel
  float* dist;               // (P, N, M) raw approximate distances out of TMEM, or null
  unsigned int* stats_out;   // host: [fix-up rows, ambiguous rows, max |approx - exact| bits], or null (forces a sync)
  int skip;                  // >= 0: sweep A leaves out every skip-th key tile (0 / 1: none)
  int ga;                    // 18 | 6: keys per group of the sweep-A list (0: automatic)
};
int launch_knn_tc(const KnnWorkspace& w, void* extra_ws, const void* x, int64_t x_sb, int64_t x_sn, int dtype, int G,
                  const float* relpos, const SepBias& sep, int32_t* idx_out, int P, int N, int M, int D, int k,
                  int dilation, const KnnDebug* dbg, cudaStream_t stream);

// cudaFuncSetAttribute is per device: run `f` once per (call site, device)
template <class F>
inline void configure_once_per_device(std::atomic<uint64_t>& done, F&& f) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (!(done.load(std::memory_order_acquire) & bit)) {
    f();
    done.fetch_or(bit, std::memory_order_release);
  }
}

}  // namespace gkg

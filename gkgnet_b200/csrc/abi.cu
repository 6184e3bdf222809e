// C-ABI glue: error reporting, launch accounting, gkg_knn_graph dispatch.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "knn_tc.cuh"

namespace gkg {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace gkg

using namespace gkg;

extern "C" int gkg_abi_version(void) { return GKG_ABI_VERSION; }
extern "C" const char* gkg_last_error(void) { return g_err; }
extern "C" uint64_t gkg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static int resolve_algo(int algo, int N, int M, int D, int k, int dilation, int dtype) {
  if (algo == GKG_KNN_AUTO)
    return knn_tc_preferred(N, M, D, k, dilation, dtype) ? GKG_KNN_TCGEN05 : GKG_KNN_EXACT_FP32;
  return algo;
}

extern "C" size_t gkg_knn_workspace_bytes(int B, int G, int N, int M, int D, int k, int dilation,
                                          int self_keys, int dtype, int algo) {
  if (B <= 0 || G <= 0 || N <= 0 || D <= 0) return 256;
  if (self_keys) M = N;
  const int P = B * G;
  const bool tc = resolve_algo(algo, N, M, D, k, dilation, dtype) == GKG_KNN_TCGEN05;
  size_t bytes = carve_knn_workspace(nullptr, P, N, M, D, self_keys != 0, tc).bytes;
  if (tc) bytes += knn_tc_workspace_bytes(P, N, M, D, k, dilation, dtype);
  return bytes + 256;
}

static int check_knn_common(int B, int G, int N, int M, int D, int k, int dilation, bool self_keys, int dtype,
                            int& algo, void* workspace, size_t workspace_bytes) {
  GKG_CHECK_ARG(B >= 0 && G > 0 && N >= 0 && D > 0 && k > 0 && dilation > 0,
                "knn_graph: bad shape B=%d G=%d N=%d D=%d k=%d dilation=%d", B, G, N, D, k, dilation);
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "knn_graph: bad dtype %d", dtype);
  GKG_CHECK_ARG(D <= 640, "knn_graph: D=%d > 640 channels per group is not supported", D);
  if ((long long)B * N == 0) return GKG_OK;
  // torch.topk raises when k*dilation exceeds the row length (reference: size < 192 fails)
  GKG_CHECK_ARG(k * dilation <= M, "knn_graph: k*dilation=%d exceeds the %d keys", k * dilation, M);
  GKG_CHECK_ARG(workspace != nullptr, "knn_graph: null workspace");
  GKG_CHECK_ARG(((uintptr_t)workspace % 256) == 0, "knn_graph: workspace must be 256-byte aligned");
  algo = resolve_algo(algo, N, M, D, k, dilation, dtype);
  GKG_CHECK_ARG(algo == GKG_KNN_EXACT_FP32 || algo == GKG_KNN_TCGEN05, "knn_graph: bad algo %d", algo);
  if (algo == GKG_KNN_TCGEN05)
    GKG_CHECK_ARG(knn_tc_supported(N, M, D, k, dilation, dtype),
                  "knn_graph: tcgen05 path does not support N=%d M=%d D=%d k*d=%d dtype=%d", N, M, D, k * dilation, dtype);
  const size_t need = gkg_knn_workspace_bytes(B, G, N, M, D, k, dilation, self_keys, dtype, algo) - 256;
  if (workspace_bytes < need) {
    set_error("knn_graph: workspace %zu < %zu bytes", workspace_bytes, need);
    return GKG_EWORKSPACE;
  }
  return GKG_OK;
}

extern "C" int gkg_knn_prepare(const void* x, int64_t x_sb, int64_t x_sn, const void* y, int64_t y_sb,
                               int64_t y_sn, int B, int G, int N, int M, int D, int k, int dilation,
                               int dtype, int algo, void* workspace, size_t workspace_bytes,
                               gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool self_keys = (y == nullptr);
  if (self_keys) M = N;
  int rc = check_knn_common(B, G, N, M, D, k, dilation, self_keys, dtype, algo, workspace, workspace_bytes);
  if (rc != GKG_OK || (long long)B * N == 0) return rc;
  GKG_CHECK_ARG(x != nullptr, "knn_prepare: null pointer");
  const int P = B * G;
  KnnWorkspace w = carve_knn_workspace(workspace, P, N, M, D, self_keys, algo == GKG_KNN_TCGEN05);
  if (algo == GKG_KNN_TCGEN05)   // fused: normalise + fp16 conversion + operand layout in one pass
    return launch_knn_tc_prepare(w, static_cast<char*>(workspace) + w.bytes, x, x_sb, x_sn, y, y_sb, y_sn,
                                 dtype, P, G, N, M, D, k, dilation, self_keys, stream);
  rc = launch_knn_prepare(x, x_sb, x_sn, dtype, w.xhat, w.xsq, B, G, N, D, stream);
  if (rc != GKG_OK) return rc;
  if (!self_keys) {
    rc = launch_knn_prepare(y, y_sb, y_sn, dtype, w.yhat, w.ysq, B, G, M, D, stream);
    if (rc != GKG_OK) return rc;
  }
  return GKG_OK;
}

static int knn_select_impl(const void* x, int64_t x_sb, int64_t x_sn, const float* relpos, const float* relpos_sep_a,
                           const float* relpos_sep_b, int sep_grid_w, int sep_kw, int32_t* idx_out, int B, int G,
                           int N, int M, int D, int k, int dilation, int self_keys_, int dtype, int algo,
                           void* workspace, size_t workspace_bytes, const KnnDebug* dbg, cudaStream_t stream) {
  const bool self_keys = self_keys_ != 0;
  if (self_keys) M = N;
  int rc = check_knn_common(B, G, N, M, D, k, dilation, self_keys, dtype, algo, workspace, workspace_bytes);
  if (rc != GKG_OK || (long long)B * N == 0) return rc;
  GKG_CHECK_ARG(idx_out != nullptr && x != nullptr, "knn_select: null pointer");
  const int P = B * G;
  KnnWorkspace w = carve_knn_workspace(workspace, P, N, M, D, self_keys, algo == GKG_KNN_TCGEN05);
  if (algo == GKG_KNN_TCGEN05) {
    SepBias sep{relpos_sep_a, relpos_sep_b, sep_grid_w, sep_kw};
    return launch_knn_tc(w, static_cast<char*>(workspace) + w.bytes, x, x_sb, x_sn, dtype, G, relpos, sep, idx_out,
                         P, N, M, D, k, dilation, dbg, stream);
  }
  return launch_knn_exact(w, relpos, idx_out, P, N, M, D, k, dilation, stream);
}

extern "C" int gkg_knn_select(const void* x, int64_t x_sb, int64_t x_sn, const float* relpos,
                              const float* relpos_sep_a, const float* relpos_sep_b, int sep_grid_w, int sep_kw,
                              int32_t* idx_out, int B, int G, int N, int M, int D, int k, int dilation,
                              int self_keys, int dtype, int algo, void* workspace, size_t workspace_bytes,
                              gkg_stream_t stream) {
  return knn_select_impl(x, x_sb, x_sn, relpos, relpos_sep_a, relpos_sep_b, sep_grid_w, sep_kw, idx_out, B, G, N, M,
                         D, k, dilation, self_keys, dtype, algo, workspace, workspace_bytes, nullptr,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int gkg_knn_graph(const void* x, int64_t x_sb, int64_t x_sn, const void* y, int64_t y_sb,
                             int64_t y_sn, const float* relpos, const float* relpos_sep_a,
                             const float* relpos_sep_b, int sep_grid_w, int sep_kw, int32_t* idx_out,
                             int B, int G, int N, int M, int D, int k, int dilation, int dtype, int algo,
                             void* workspace, size_t workspace_bytes, gkg_stream_t stream) {
  int rc = gkg_knn_prepare(x, x_sb, x_sn, y, y_sb, y_sn, B, G, N, M, D, k, dilation, dtype, algo,
                           workspace, workspace_bytes, stream);
  if (rc != GKG_OK) return rc;
  return gkg_knn_select(x, x_sb, x_sn, relpos, relpos_sep_a, relpos_sep_b, sep_grid_w, sep_kw, idx_out, B, G, N, M,
                        D, k, dilation, y == nullptr, dtype, algo, workspace, workspace_bytes, stream);
}

// Test hook (declared in include/gkg_abi.h under "testing"): gkg_knn_graph with per-call debug options; nothing
// here is process state.  debug_flags: 1 = re-rank every row exactly, 3 = every row through the brute-force
// fix-up kernel; dbg_dist: (B*G, N, M) fp32 device buffer that receives the raw approximate distances, or NULL;
// stats_out: 3 host words [fix-up rows, ambiguous rows, bits of the largest |approx - exact|] (synchronises).
extern "C" int gkg_knn_graph_debug(const void* x, int64_t x_sb, int64_t x_sn, const void* y, int64_t y_sb,
                                   int64_t y_sn, const float* relpos, const float* relpos_sep_a,
                                   const float* relpos_sep_b, int sep_grid_w, int sep_kw, int32_t* idx_out,
                                   int B, int G, int N, int M, int D, int k, int dilation, int dtype, int algo,
                                   void* workspace, size_t workspace_bytes, gkg_stream_t stream,
                                   int debug_flags, float* dbg_dist, unsigned int* stats_out, int skip, int ga) {
  int rc = gkg_knn_prepare(x, x_sb, x_sn, y, y_sb, y_sn, B, G, N, M, D, k, dilation, dtype, algo,
                           workspace, workspace_bytes, stream);
  if (rc != GKG_OK) return rc;
  KnnDebug dbg{debug_flags, dbg_dist, stats_out, skip, ga};
  return knn_select_impl(x, x_sb, x_sn, relpos, relpos_sep_a, relpos_sep_b, sep_grid_w, sep_kw, idx_out, B, G, N, M,
                         D, k, dilation, y == nullptr, dtype, algo, workspace, workspace_bytes, &dbg,
                         static_cast<cudaStream_t>(stream));
}

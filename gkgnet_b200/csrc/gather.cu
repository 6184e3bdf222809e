// Neighbour gather and neighbour sum on the token-major layout: the `batched_index_select` of the reference
// (torch_nn.py:84-105: x_j[b, c, n, j] = y[b, c, idx[b, n, j]]) for the graph convolutions that need the gathered
// rows themselves (EdgeConv2d, GraphSAGE, GraphAtten: torch_vertex.py:16-131), and its sum over the neighbours for
// GINConv2d (:134-150).  The max-relative convolution GKGNet uses never materialises this tensor (mr_aggregate.cu).
//   out[b, n, j, c] = y[b, idx[b*G + c / D, n, j], c]            (gather)
//   out[b, n, c]    = sum_j y[b, idx[b*G + c / D, n, j], c]      (sum, fp32 accumulation, one rounding)
// Backward of both: scatter-add into an fp32 (B, M, C) accumulator (atomics), like index_put_(accumulate=True).
#include "common.cuh"

namespace gkg {
namespace {

template <typename T, bool SUM>
__global__ void neighbor_gather_fwd_kernel(const T* __restrict__ y, int64_t y_sb, int64_t y_sn, const int32_t* __restrict__ idx,
                                           T* __restrict__ out, long long total, int G, int N, int D, int k) {
  const int C = G * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long t = e / C;
    int j = 0;
    if (!SUM) { j = (int)(t % k); t /= k; }
    const int n = (int)(t % N);
    const long long b = t / N;
    const int32_t* row = idx + ((b * G + c / D) * N + n) * (long long)k;
    if (SUM) {
      float acc = 0.f;
      for (int jj = 0; jj < k; ++jj) acc += to_f32<T>(y[b * y_sb + (long long)row[jj] * y_sn + c]);
      out[e] = from_f32<T>(acc);
    } else {
      out[e] = y[b * y_sb + (long long)row[j] * y_sn + c];
    }
  }
}

template <typename T, bool SUM>
__global__ void neighbor_gather_bwd_kernel(const T* __restrict__ grad, const int32_t* __restrict__ idx, float* __restrict__ gy,
                                           long long total, int G, int N, int M, int D, int k) {
  const int C = G * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long t = e / C;
    int j = 0;
    if (!SUM) { j = (int)(t % k); t /= k; }
    const int n = (int)(t % N);
    const long long b = t / N;
    const int32_t* row = idx + ((b * G + c / D) * N + n) * (long long)k;
    const float g = to_f32<T>(grad[e]);
    if (SUM) {
      for (int jj = 0; jj < k; ++jj) atomicAdd(gy + (b * M + row[jj]) * C + c, g);
    } else {
      atomicAdd(gy + (b * M + row[j]) * C + c, g);
    }
  }
}

template <bool SUM>
int launch_fwd(const void* y, int64_t y_sb, int64_t y_sn, const int32_t* idx, void* out, int B, int G, int N, int M, int D,
               int k, int dtype, cudaStream_t stream) {
  (void)M;
  const long long total = (long long)B * N * (SUM ? 1 : k) * G * D;
  if (total == 0) return GKG_OK;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == GKG_F32)
    neighbor_gather_fwd_kernel<float, SUM><<<blocks, 256, 0, stream>>>(static_cast<const float*>(y), y_sb, y_sn, idx,
                                                                       static_cast<float*>(out), total, G, N, D, k);
  else
    neighbor_gather_fwd_kernel<__nv_bfloat16, SUM><<<blocks, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(y), y_sb, y_sn, idx, static_cast<__nv_bfloat16*>(out), total, G, N, D, k);
  GKG_CHECK_LAUNCH("neighbor_gather_fwd_kernel");
  return GKG_OK;
}

template <bool SUM>
int launch_bwd(const void* grad, const int32_t* idx, float* gy, int B, int G, int N, int M, int D, int k, int dtype,
               cudaStream_t stream) {
  const long long total = (long long)B * N * (SUM ? 1 : k) * G * D;
  if (total == 0) return GKG_OK;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == GKG_F32)
    neighbor_gather_bwd_kernel<float, SUM><<<blocks, 256, 0, stream>>>(static_cast<const float*>(grad), idx, gy, total, G, N, M, D, k);
  else
    neighbor_gather_bwd_kernel<__nv_bfloat16, SUM><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(grad), idx, gy,
                                                                               total, G, N, M, D, k);
  GKG_CHECK_LAUNCH("neighbor_gather_bwd_kernel");
  return GKG_OK;
}

int check(const void* a, const void* b, const void* c, int B, int G, int N, int M, int D, int k, int dtype, const char* what) {
  GKG_CHECK_ARG(B >= 0 && G > 0 && N >= 0 && M > 0 && D > 0 && k > 0, "%s: bad shape", what);
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "%s: bad dtype %d", what, dtype);
  if ((long long)B * N == 0) return GKG_OK;
  GKG_CHECK_ARG(a && b && c, "%s: null pointer", what);
  return GKG_OK;
}

}  // namespace
}  // namespace gkg

using namespace gkg;

extern "C" int gkg_neighbor_gather_fwd(const void* y, int64_t y_sb, int64_t y_sn, const int32_t* idx, void* out, int B, int G,
                                       int N, int M, int D, int k, int dtype, gkg_stream_t stream) {
  int rc = check(y, idx, out, B, G, N, M, D, k, dtype, "neighbor_gather_fwd");
  return rc != GKG_OK ? rc : launch_fwd<false>(y, y_sb, y_sn, idx, out, B, G, N, M, D, k, dtype, static_cast<cudaStream_t>(stream));
}
extern "C" int gkg_neighbor_gather_bwd(const void* grad, const int32_t* idx, float* grad_y_accum, int B, int G, int N, int M,
                                       int D, int k, int dtype, gkg_stream_t stream) {
  int rc = check(grad, idx, grad_y_accum, B, G, N, M, D, k, dtype, "neighbor_gather_bwd");
  return rc != GKG_OK ? rc : launch_bwd<false>(grad, idx, grad_y_accum, B, G, N, M, D, k, dtype, static_cast<cudaStream_t>(stream));
}
extern "C" int gkg_neighbor_sum_fwd(const void* y, int64_t y_sb, int64_t y_sn, const int32_t* idx, void* out, int B, int G, int N,
                                    int M, int D, int k, int dtype, gkg_stream_t stream) {
  int rc = check(y, idx, out, B, G, N, M, D, k, dtype, "neighbor_sum_fwd");
  return rc != GKG_OK ? rc : launch_fwd<true>(y, y_sb, y_sn, idx, out, B, G, N, M, D, k, dtype, static_cast<cudaStream_t>(stream));
}
extern "C" int gkg_neighbor_sum_bwd(const void* grad, const int32_t* idx, float* grad_y_accum, int B, int G, int N, int M, int D,
                                    int k, int dtype, gkg_stream_t stream) {
  int rc = check(grad, idx, grad_y_accum, B, G, N, M, D, k, dtype, "neighbor_sum_bwd");
  return rc != GKG_OK ? rc : launch_bwd<true>(grad, idx, grad_y_accum, B, G, N, M, D, k, dtype, static_cast<cudaStream_t>(stream));
}

// kNN operand preparation: per-group L2 normalisation of node features.
//
// Follows F.normalize(x, p=2, dim=1) at torch_edge.py:167-168,173 of the reference:
//   xhat = x / max(||x||_2, 1e-12)   over the D channels of one group,
// and the squared norms the distance formula adds back (torch_edge.py:49-50).
// Input is token-major (b, n, c) with c = g*D + d; output is (P = B*G, N, D) fp32.
#include "common.cuh"

namespace gkg {

template <typename T>
__global__ void __launch_bounds__(256)
knn_prepare_kernel(const T* __restrict__ feat, int64_t stride_b, int64_t stride_n,
                   float* __restrict__ hat, float* __restrict__ sq, int G, int N, int D,
                   long long rows) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= rows) return;
  const int n = (int)(warp % N);
  const long long p = warp / N;
  const int g = (int)(p % G);
  const long long b = p / G;
  const T* src = feat + b * stride_b + (long long)n * stride_n + (long long)g * D;
  float* dst = hat + warp * D;

  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    float v = to_f32<T>(src[d]);
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  float s2 = 0.f;
  for (int d = lane; d < D; d += 32) {
    float v = to_f32<T>(src[d]) / denom;  // true division, like aten::div
    dst[d] = v;
    s2 = fmaf(v, v, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if (lane == 0) sq[warp] = s2;
}

int launch_knn_prepare(const void* feat, int64_t stride_b, int64_t stride_n, int dtype, float* hat,
                       float* sq, int B, int G, int N, int D, cudaStream_t stream) {
  const long long rows = (long long)B * G * N;
  if (rows == 0) return GKG_OK;
  const int warps_per_block = 8;
  const long long blocks = (rows + warps_per_block - 1) / warps_per_block;
  GKG_CHECK_ARG(blocks < 0x7fffffffLL, "knn_prepare: too many rows (%lld)", rows);
  if (dtype == GKG_F32) {
    knn_prepare_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(
        static_cast<const float*>(feat), stride_b, stride_n, hat, sq, G, N, D, rows);
  } else {
    knn_prepare_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(feat), stride_b, stride_n, hat, sq, G, N, D, rows);
  }
  GKG_CHECK_LAUNCH("knn_prepare");
  return GKG_OK;
}

}  // namespace gkg

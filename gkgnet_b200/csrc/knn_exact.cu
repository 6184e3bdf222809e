// Exact fp32 brute-force kNN (CUDA cores).  One thread per query row, keys staged in
// shared memory.  Implements, in the reference's association order,
//   dist = (|xh|^2 + (-2 * xh.yh)) + |yh|^2  [+ relative_pos]     torch_edge.py:48-51,102-103
//   topk(-dist, k*d) sorted ascending, keep ranks 0, d, 2d, ...   torch_edge.py:104,148
// This is the always-available path (any D <= 640, any k*d <= 64) and the exactness
// yard-stick for the tcgen05 kernel; it never materialises the N x M matrix either.
#include "knn_tc.cuh"

namespace gkg {

constexpr int kExactRowsWide = 128;   // query rows per CTA (one per thread); 32 for D > 320 (shared memory)
constexpr int kExactKeys = 32;    // keys per shared-memory tile
constexpr int kExactMaxKD = 64;   // max k*dilation

template <int kExactRows>
__global__ void __launch_bounds__(kExactRows)
knn_exact_kernel(const float* __restrict__ xhat, const float* __restrict__ xsq,
                 const float* __restrict__ yhat, const float* __restrict__ ysq,
                 const float* __restrict__ relpos, int32_t* __restrict__ idx_out, int N, int M,
                 int D, int k, int dilation) {
  extern __shared__ float4 smem4[];
  const int D4 = (D + 3) >> 2;
  float4* xs = smem4;                         // [D4][128]
  float4* ys = smem4 + D4 * kExactRows;       // [kExactKeys][D4]
  float* ysq_s = reinterpret_cast<float*>(ys + kExactKeys * D4);  // [kExactKeys]

  const int t = threadIdx.x;
  const long long p = blockIdx.y;
  const int r0 = blockIdx.x * kExactRows;
  const int n = r0 + t;
  const bool valid = n < N;
  const int kd = k * dilation;

  // stage the query tile, transposed to [d/4][row] so that lanes read consecutive float4s
  {
    float* xs_f = reinterpret_cast<float*>(xs);
    const int rows_here = min(kExactRows, N - r0);
    const float* src = xhat + (p * N + r0) * (long long)D;
    const int Dp = D4 * 4;
    for (int i = t; i < kExactRows * Dp; i += kExactRows) {
      const int row = i / Dp, d = i - row * Dp;
      float v = 0.f;
      if (row < rows_here && d < D) v = src[(long long)row * D + d];
      xs_f[((d >> 2) * kExactRows + row) * 4 + (d & 3)] = v;
    }
  }
  const float my_xsq = valid ? xsq[p * N + n] : 0.f;
  const float* my_rel = (relpos != nullptr && valid) ? relpos + (long long)n * M : nullptr;

  float lv[kExactMaxKD];
  int li[kExactMaxKD];
#pragma unroll 1
  for (int j = 0; j < kd; ++j) { lv[j] = INFINITY; li[j] = 0; }
  float tau = INFINITY;

  const float* ysrc = yhat + p * (long long)M * D;
  const float* ysq_src = ysq + p * (long long)M;

  for (int m0 = 0; m0 < M; m0 += kExactKeys) {
    __syncthreads();
    {
      float* ys_f = reinterpret_cast<float*>(ys);
      const int keys_here = min(kExactKeys, M - m0);
      const int Dp = D4 * 4;
      for (int i = t; i < kExactKeys * Dp; i += kExactRows) {
        const int kk = i / Dp, d = i - kk * Dp;
        float v = 0.f;
        if (kk < keys_here && d < D) v = ysrc[(long long)(m0 + kk) * D + d];
        ys_f[i] = v;
      }
      if (t < kExactKeys) ysq_s[t] = (t < keys_here) ? ysq_src[m0 + t] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int kk = 0; kk < kExactKeys; kk += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int d4 = 0; d4 < D4; ++d4) {
        const float4 xv = xs[d4 * kExactRows + t];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 yv = ys[(kk + j) * D4 + d4];
          acc[j] = fmaf(xv.x, yv.x, acc[j]);
          acc[j] = fmaf(xv.y, yv.y, acc[j]);
          acc[j] = fmaf(xv.z, yv.z, acc[j]);
          acc[j] = fmaf(xv.w, yv.w, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int m = m0 + kk + j;
        if (valid && m < M) {
          float v = (my_xsq + (-2.f * acc[j])) + ysq_s[kk + j];
          if (my_rel != nullptr) v += my_rel[m];
          if (v < tau) {
            int q = kd - 1;
            while (q > 0 && lv[q - 1] > v) {
              lv[q] = lv[q - 1];
              li[q] = li[q - 1];
              --q;
            }
            lv[q] = v;
            li[q] = m;
            tau = lv[kd - 1];
          }
        }
      }
    }
  }
  if (valid) {
    int32_t* out = idx_out + (p * N + n) * (long long)k;
    for (int j = 0; j < k; ++j) out[j] = li[j * dilation];
  }
}

int launch_knn_exact(const KnnWorkspace& w, const float* relpos, int32_t* idx_out, int P, int N,
                     int M, int D, int k, int dilation, cudaStream_t stream) {
  GKG_CHECK_ARG(k * dilation <= kExactMaxKD, "knn_exact: k*dilation=%d > %d", k * dilation,
                kExactMaxKD);
  GKG_CHECK_ARG(P <= 65535, "knn_exact: B*G=%d > 65535", P);
  const int D4 = (D + 3) / 4;
  const int rows_per_cta = D <= 320 ? kExactRowsWide : 32;
  const size_t smem = sizeof(float4) * ((size_t)D4 * rows_per_cta + (size_t)kExactKeys * D4) +
                      sizeof(float) * kExactKeys;
  GKG_CHECK_ARG(smem <= 227 * 1024, "knn_exact: D=%d needs %zu B of shared memory", D, smem);
  auto kern = D <= 320 ? knn_exact_kernel<kExactRowsWide> : knn_exact_kernel<32>;
  {
    static std::atomic<uint64_t> configured[2];
    cudaError_t e = cudaSuccess;
    configure_once_per_device(configured[D <= 320 ? 0 : 1], [&] {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (e != cudaSuccess) {
      set_error("knn_exact: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return GKG_ECUDA;
    }
  }
  dim3 grid((N + rows_per_cta - 1) / rows_per_cta, P);
  kern<<<grid, rows_per_cta, smem, stream>>>(w.xhat, w.xsq, w.yhat, w.ysq, relpos, idx_out, N, M, D, k, dilation);
  GKG_CHECK_LAUNCH("knn_exact");
  return GKG_OK;
}

}  // namespace gkg

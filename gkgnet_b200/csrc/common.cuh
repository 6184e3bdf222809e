// Shared helpers for libgkg_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gkg_abi.h"

namespace gkg {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GKG_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      gkg::set_error(__VA_ARGS__);          \
      return GKG_EINVAL;                    \
    }                                       \
  } while (0)

#define GKG_CHECK_LAUNCH(name)                                                    \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      gkg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return GKG_ECUDA;                                                           \
    }                                                                             \
    gkg::count_launch();                                                          \
  } while (0)

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// ---- kNN workspace carve-up (shared by host wrappers) -------------------------------
struct KnnWorkspace {
  float* xhat;   // (P, N, D) normalised queries, fp32 (exact path only: the tensor-core path rebuilds the few rows it needs)
  float* xsq;    // (P, N)    |xhat|^2
  float* yhat;   // (P, M, D) normalised keys (== xhat when self on the exact path)
  float* ysq;    // (P, M)
  size_t bytes;
};

inline KnnWorkspace carve_knn_workspace(void* base, int P, int N, int M, int D, bool self_keys, bool tc) {
  KnnWorkspace w;
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t n) { void* p = b ? b + off : nullptr; off += align_up(n, 256); return p; };
  if (tc) {
    w.xhat = nullptr;
    w.xsq = nullptr;
    w.yhat = static_cast<float*>(take(sizeof(float) * (size_t)P * M * D));
    w.ysq = static_cast<float*>(take(sizeof(float) * (size_t)P * M));
    w.bytes = off;
    return w;
  }
  w.xhat = static_cast<float*>(take(sizeof(float) * (size_t)P * N * D));
  w.xsq = static_cast<float*>(take(sizeof(float) * (size_t)P * N));
  if (self_keys) {
    w.yhat = w.xhat;
    w.ysq = w.xsq;
  } else {
    w.yhat = static_cast<float*>(take(sizeof(float) * (size_t)P * M * D));
    w.ysq = static_cast<float*>(take(sizeof(float) * (size_t)P * M));
  }
  w.bytes = off;
  return w;
}

// kernels' host launchers (defined in the .cu files)
int launch_knn_prepare(const void* feat, int64_t stride_b, int64_t stride_n, int dtype, float* hat,
                       float* sq, int B, int G, int N, int D, cudaStream_t stream);
int launch_knn_exact(const KnnWorkspace& w, const float* relpos, int32_t* idx_out, int P, int N,
                     int M, int D, int k, int dilation, cudaStream_t stream);

}  // namespace gkg

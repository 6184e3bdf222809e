// Key pooling of the dynamic graph convolution: y = avg_pool2d(x, r, r) of
// DyGraphConv2d(MultiGroup).forward (torch_vertex.py:194-196 / :221-223), on the token-major
// (B, H*W, C) layout of this library.  HBM bound: every input byte is read once with 128-bit loads
// (the r*r loads of a thread are independent), the output is 1/r^2 of the input.
//
// Numerics follow ATen's channels-last kernel: the window is summed in fp32 in row-major order
// (kh outer, kw inner) and divided by r*r once, then rounded to the feature dtype.  Windows are
// always complete (floor mode, no padding); rows / columns past floor(H/r)*r are ignored and get a
// zero gradient, as in the reference.
#include "common.cuh"

namespace gkg {
namespace {

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) PoolPack {
  T v[VEC];
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
pool_keys_fwd_kernel(const T* __restrict__ x, int64_t x_sb, int64_t x_sn, T* __restrict__ y, int H, int W, int C,
                     int r, int Ho, int Wo, long long total) {
  using P = PoolPack<T, VEC>;
  const int cpn = C / VEC;
  const float div = (float)(r * r);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % cpn);
    long long t = i / cpn;
    const int ow = (int)(t % Wo);
    t /= Wo;
    const int oh = (int)(t % Ho);
    const long long b = t / Ho;
    const T* base = x + b * x_sb + ((long long)(oh * r) * W + ow * r) * x_sn + cc * VEC;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int kh = 0; kh < r; ++kh) {
#pragma unroll 4
      for (int kw = 0; kw < r; ++kw) {
        const P v = *reinterpret_cast<const P*>(base + ((long long)kh * W + kw) * x_sn);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += to_f32<T>(v.v[e]);
      }
    }
    P o;
#pragma unroll
    for (int e = 0; e < VEC; ++e) o.v[e] = from_f32<T>(acc[e] / div);
    *reinterpret_cast<P*>(y + ((b * Ho + oh) * (long long)Wo + ow) * C + cc * VEC) = o;
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
pool_keys_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int H, int W, int C, int r, int Ho, int Wo,
                     long long total) {
  using P = PoolPack<T, VEC>;
  const int cpn = C / VEC;
  const float div = (float)(r * r);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % cpn);
    long long t = i / cpn;
    const int iw = (int)(t % W);
    t /= W;
    const int ih = (int)(t % H);
    const long long b = t / H;
    const int oh = ih / r, ow = iw / r;
    P o;
    if (oh < Ho && ow < Wo) {
      const P g = *reinterpret_cast<const P*>(gy + ((b * Ho + oh) * (long long)Wo + ow) * C + cc * VEC);
#pragma unroll
      for (int e = 0; e < VEC; ++e) o.v[e] = from_f32<T>(to_f32<T>(g.v[e]) / div);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o.v[e] = from_f32<T>(0.f);
    }
    *reinterpret_cast<P*>(gx + ((b * H + ih) * (long long)W + iw) * C + cc * VEC) = o;
  }
}

int pool_vec(int dtype, int C, const void* a, const void* b, int64_t s0, int64_t s1) {
  const int es = dtype == GKG_F32 ? 4 : 2;
  for (int vec = 16 / es; vec > 1; vec >>= 1) {
    const uintptr_t bytes = (uintptr_t)vec * es;
    if (C % vec == 0 && ((uintptr_t)a % bytes) == 0 && ((uintptr_t)b % bytes) == 0 && s0 % vec == 0 && s1 % vec == 0)
      return vec;
  }
  return 1;
}

unsigned pool_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = 148LL * 8 * 8;
  return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace
}  // namespace gkg

using namespace gkg;

extern "C" int gkg_pool_keys_fwd(const void* x, int64_t x_sb, int64_t x_sn, void* y, int B, int H, int W, int C,
                                 int r, int dtype, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "pool_keys_fwd: bad dtype %d", dtype);
  GKG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && r > 0, "pool_keys_fwd: bad shape B=%d H=%d W=%d C=%d r=%d", B,
                H, W, C, r);
  GKG_CHECK_ARG(x && y, "pool_keys_fwd: null pointer");
  const int Ho = H / r, Wo = W / r;
  GKG_CHECK_ARG(Ho > 0 && Wo > 0, "pool_keys_fwd: window %d larger than the %dx%d map", r, H, W);
  if (B == 0) return GKG_OK;
  const int vec = pool_vec(dtype, C, x, y, x_sb, x_sn);
  const long long total = (long long)B * Ho * Wo * (C / vec);
  const unsigned grid = pool_grid(total);
#define LAUNCH(T, V)                                                                                       \
  pool_keys_fwd_kernel<T, V><<<grid, 256, 0, stream>>>(static_cast<const T*>(x), x_sb, x_sn, static_cast<T*>(y), H, \
                                                       W, C, r, Ho, Wo, total)
  if (dtype == GKG_F32) {
    if (vec == 4) LAUNCH(float, 4); else if (vec == 2) LAUNCH(float, 2); else LAUNCH(float, 1);
  } else {
    if (vec == 8) LAUNCH(__nv_bfloat16, 8); else if (vec == 4) LAUNCH(__nv_bfloat16, 4);
    else if (vec == 2) LAUNCH(__nv_bfloat16, 2); else LAUNCH(__nv_bfloat16, 1);
  }
#undef LAUNCH
  GKG_CHECK_LAUNCH("pool_keys_fwd");
  return GKG_OK;
}

extern "C" int gkg_pool_keys_bwd(const void* grad_y, void* grad_x, int B, int H, int W, int C, int r, int dtype,
                                 gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "pool_keys_bwd: bad dtype %d", dtype);
  GKG_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && r > 0, "pool_keys_bwd: bad shape B=%d H=%d W=%d C=%d r=%d", B,
                H, W, C, r);
  GKG_CHECK_ARG(grad_y && grad_x, "pool_keys_bwd: null pointer");
  const int Ho = H / r, Wo = W / r;
  GKG_CHECK_ARG(Ho > 0 && Wo > 0, "pool_keys_bwd: window %d larger than the %dx%d map", r, H, W);
  if (B == 0) return GKG_OK;
  const int vec = pool_vec(dtype, C, grad_y, grad_x, 0, 0);
  const long long total = (long long)B * H * W * (C / vec);
  const unsigned grid = pool_grid(total);
#define LAUNCH(T, V)                                                                                              \
  pool_keys_bwd_kernel<T, V><<<grid, 256, 0, stream>>>(static_cast<const T*>(grad_y), static_cast<T*>(grad_x), H, W, C, \
                                                       r, Ho, Wo, total)
  if (dtype == GKG_F32) {
    if (vec == 4) LAUNCH(float, 4); else if (vec == 2) LAUNCH(float, 2); else LAUNCH(float, 1);
  } else {
    if (vec == 8) LAUNCH(__nv_bfloat16, 8); else if (vec == 4) LAUNCH(__nv_bfloat16, 4);
    else if (vec == 2) LAUNCH(__nv_bfloat16, 2); else LAUNCH(__nv_bfloat16, 1);
  }
#undef LAUNCH
  GKG_CHECK_LAUNCH("pool_keys_bwd");
  return GKG_OK;
}

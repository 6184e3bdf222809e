// Max-relative graph aggregation (forward + backward), memory-bound gather kernels.
//
// Forward replaces the reference's two batched_index_select gathers (torch_nn.py:84-105),
// the subtraction + max over k (torch_vertex.py:54) and the channel-interleaving cat
// (torch_vertex.py:57-61) with ONE pass: nothing of size (B, C, N, k) is materialised.
//   out[b, n, 2c]   = x[b, n, c]
//   out[b, n, 2c+1] = max_j y[b, idx[p, n, j], c] - x[b, n, c]        (p = b*G + c / D)
// max_j(x_j - x_i) == max_j(x_j) - x_i bit-exactly (rounding is monotone), so the centre
// row is subtracted once.  Backward routes grad to the arg-max neighbour (what autograd
// derives from torch.max + index_put_(accumulate=True)).
//
// Layout: token-major (b, n, c): one thread owns a 16-byte channel chunk of one node, so
// x reads, idx reads, out writes are coalesced and every neighbour row is fetched with
// 128-bit loads (served from L2: the key set of one image is <= a few hundred KB).
#include "common.cuh"

namespace gkg {

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
  T v[VEC];
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
mr_aggregate_fwd_kernel(const T* __restrict__ x, int64_t x_sb, int64_t x_sn,
                        const T* __restrict__ y, int64_t y_sb, int64_t y_sn,
                        const int32_t* __restrict__ idx, T* __restrict__ out,
                        uint8_t* __restrict__ argmax, int G, int N, int D, int k,
                        long long total_chunks) {
  using P = Pack<T, VEC>;
  const int C = G * D;
  const int chunks_per_node = C / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % chunks_per_node);
    const long long bn = i / chunks_per_node;
    const int n = (int)(bn % N);
    const long long b = bn / N;
    const int c0 = cc * VEC;
    const int g = c0 / D;
    const int32_t* ip = idx + ((b * G + g) * N + n) * (long long)k;
    const P xv = *reinterpret_cast<const P*>(x + b * x_sb + (long long)n * x_sn + c0);
    const T* ybase = y + b * y_sb + c0;

    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll 3
    for (int j = 0; j < k; ++j) {
      const int m = __ldg(ip + j);
      const P yv = *reinterpret_cast<const P*>(ybase + (long long)m * y_sn);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float f = to_f32<T>(yv.v[e]);
        if (f > best[e]) { best[e] = f; arg[e] = j; }
      }
    }
    Pack<T, 2 * VEC> o;  // 2*VEC interleaved outputs [x_c, m_c, ...]
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      o.v[2 * e] = xv.v[e];
      o.v[2 * e + 1] = from_f32<T>(best[e] - to_f32<T>(xv.v[e]));
    }
    T* op = out + (bn * C + c0) * 2;
    if constexpr (sizeof(T) * 2 * VEC <= 16) {
      *reinterpret_cast<Pack<T, 2 * VEC>*>(op) = o;
    } else {
      const P* halves = reinterpret_cast<const P*>(&o);
      reinterpret_cast<P*>(op)[0] = halves[0];
      reinterpret_cast<P*>(op)[1] = halves[1];
    }
    if (argmax != nullptr) {
      Pack<uint8_t, VEC> a;
#pragma unroll
      for (int e = 0; e < VEC; ++e) a.v[e] = (uint8_t)arg[e];
      *reinterpret_cast<Pack<uint8_t, VEC>*>(argmax + bn * C + c0) = a;
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
mr_aggregate_bwd_kernel(const T* __restrict__ gout, const int32_t* __restrict__ idx,
                        const uint8_t* __restrict__ argmax, T* __restrict__ gx,
                        float* __restrict__ gy, int G, int N, int M, int D, int k,
                        long long total_chunks) {
  using P = Pack<T, VEC>;
  const int C = G * D;
  const int chunks_per_node = C / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % chunks_per_node);
    const long long bn = i / chunks_per_node;
    const int n = (int)(bn % N);
    const long long b = bn / N;
    const int c0 = cc * VEC;
    const int g = c0 / D;
    const int32_t* ip = idx + ((b * G + g) * N + n) * (long long)k;
    const T* gp = gout + (bn * C + c0) * 2;
    float g_self[VEC], g_rel[VEC];
    {
      Pack<T, 2 * VEC> gi;
      if constexpr (sizeof(T) * 2 * VEC <= 16) {
        gi = *reinterpret_cast<const Pack<T, 2 * VEC>*>(gp);
      } else {
        P* halves = reinterpret_cast<P*>(&gi);
        halves[0] = reinterpret_cast<const P*>(gp)[0];
        halves[1] = reinterpret_cast<const P*>(gp)[1];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        g_self[e] = to_f32<T>(gi.v[2 * e]);
        g_rel[e] = to_f32<T>(gi.v[2 * e + 1]);
      }
    }
    const Pack<uint8_t, VEC> am = *reinterpret_cast<const Pack<uint8_t, VEC>*>(argmax + bn * C + c0);
    P o;
    float* gyb = gy + (b * M) * (long long)C + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      o.v[e] = from_f32<T>(g_self[e] - g_rel[e]);
      const int m = __ldg(ip + am.v[e]);
      atomicAdd(gyb + (long long)m * C + e, g_rel[e]);
    }
    *reinterpret_cast<P*>(gx + bn * C + c0) = o;
  }
}

static int pick_vec(int dtype, int D, int C, const void* a, const void* b, int64_t s0, int64_t s1,
                    int64_t s2, int64_t s3) {
  const int es = dtype == GKG_F32 ? 4 : 2;
  for (int vec = 16 / es; vec > 1; vec >>= 1) {
    const uintptr_t bytes = (uintptr_t)vec * es;
    bool ok = D % vec == 0 && C % vec == 0 && ((uintptr_t)a % bytes) == 0 &&
              ((uintptr_t)b % bytes) == 0 && s0 % vec == 0 && s1 % vec == 0 && s2 % vec == 0 &&
              s3 % vec == 0;
    if (ok) return vec;
  }
  return 1;
}

static unsigned grid_for(long long work_items, int block) {
  long long blocks = (work_items + block - 1) / block;
  const long long cap = 148LL * 8 * 16;  // grid-stride: a few waves of 8 CTAs/SM
  return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace gkg

using namespace gkg;

extern "C" int gkg_mr_aggregate_fwd(const void* x, int64_t x_sb, int64_t x_sn, const void* y,
                                    int64_t y_sb, int64_t y_sn, const int32_t* idx, void* out,
                                    uint8_t* argmax, int B, int G, int N, int M, int D, int k,
                                    int dtype, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "mr_aggregate_fwd: bad dtype %d", dtype);
  GKG_CHECK_ARG(B >= 0 && G > 0 && N >= 0 && D > 0 && k > 0 && k <= 255,
                "mr_aggregate_fwd: bad shape B=%d G=%d N=%d D=%d k=%d", B, G, N, D, k);
  GKG_CHECK_ARG(x && idx && out, "mr_aggregate_fwd: null pointer");
  if (y == nullptr) { y = x; y_sb = x_sb; y_sn = x_sn; M = N; }
  GKG_CHECK_ARG(M > 0 || N == 0, "mr_aggregate_fwd: no keys");
  if ((long long)B * N == 0) return GKG_OK;
  const int C = G * D;
  const int vec = pick_vec(dtype, D, C, x, y, x_sb, x_sn, y_sb, y_sn);
  const bool out_ok = ((uintptr_t)out % 16) == 0 && (argmax == nullptr || ((uintptr_t)argmax % 8) == 0);
  GKG_CHECK_ARG(out_ok, "mr_aggregate_fwd: out must be 16-byte aligned, argmax 8-byte aligned");
  const long long chunks = (long long)B * N * (C / vec);
  const unsigned grid = grid_for(chunks, 256);
#define LAUNCH(T, V)                                                                          \
  mr_aggregate_fwd_kernel<T, V><<<grid, 256, 0, stream>>>(                                    \
      static_cast<const T*>(x), x_sb, x_sn, static_cast<const T*>(y), y_sb, y_sn, idx,        \
      static_cast<T*>(out), argmax, G, N, D, k, chunks)
  if (dtype == GKG_F32) {
    if (vec == 4) LAUNCH(float, 4); else if (vec == 2) LAUNCH(float, 2); else LAUNCH(float, 1);
  } else {
    if (vec == 8) LAUNCH(__nv_bfloat16, 8); else if (vec == 4) LAUNCH(__nv_bfloat16, 4);
    else if (vec == 2) LAUNCH(__nv_bfloat16, 2); else LAUNCH(__nv_bfloat16, 1);
  }
#undef LAUNCH
  GKG_CHECK_LAUNCH("mr_aggregate_fwd");
  return GKG_OK;
}

extern "C" int gkg_mr_aggregate_bwd(const void* grad_out, const int32_t* idx, const uint8_t* argmax,
                                    void* grad_x, float* grad_y_accum, int B, int G, int N, int M,
                                    int D, int k, int dtype, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "mr_aggregate_bwd: bad dtype %d", dtype);
  GKG_CHECK_ARG(B >= 0 && G > 0 && N >= 0 && M > 0 && D > 0 && k > 0 && k <= 255,
                "mr_aggregate_bwd: bad shape B=%d G=%d N=%d M=%d D=%d k=%d", B, G, N, M, D, k);
  GKG_CHECK_ARG(grad_out && idx && argmax && grad_x && grad_y_accum, "mr_aggregate_bwd: null pointer");
  if ((long long)B * N == 0) return GKG_OK;
  const int C = G * D;
  int vec = pick_vec(dtype, D, C, grad_out, grad_x, 0, 0, 0, 0);
  while (vec > 1 && ((uintptr_t)argmax % vec) != 0) vec >>= 1;
  const long long chunks = (long long)B * N * (C / vec);
  const unsigned grid = grid_for(chunks, 256);
#define LAUNCH(T, V)                                                                           \
  mr_aggregate_bwd_kernel<T, V><<<grid, 256, 0, stream>>>(                                     \
      static_cast<const T*>(grad_out), idx, argmax, static_cast<T*>(grad_x), grad_y_accum, G, \
      N, M, D, k, chunks)
  if (dtype == GKG_F32) {
    if (vec == 4) LAUNCH(float, 4); else if (vec == 2) LAUNCH(float, 2); else LAUNCH(float, 1);
  } else {
    if (vec == 8) LAUNCH(__nv_bfloat16, 8); else if (vec == 4) LAUNCH(__nv_bfloat16, 4);
    else if (vec == 2) LAUNCH(__nv_bfloat16, 2); else LAUNCH(__nv_bfloat16, 1);
  }
#undef LAUNCH
  GKG_CHECK_LAUNCH("mr_aggregate_bwd");
  return GKG_OK;
}

// Training-mode batch-norm reductions on the token-major (rows, C) layout of this library.
//
// Reference: every conv -> norm pair of the graph blocks is `norm_layer('batch')` = (Sync)BatchNorm over
// (B, C, H, W) (torch_nn.py:32-42; BasicConv torch_nn.py:61-65, Grapher.fc1 / fc2 torch_vertex.py:290-306, FFN
// gkgnet.py:46-72): y = (x - mean_c) * rsqrt(var_c + eps) * gamma_c + beta_c with the biased batch variance, running
// statistics updated with the unbiased one.  The forward needs per-channel (mean, invstd) BEFORE the elementwise pass,
// the backward per-channel (sum dy, sum dy * (x - mean)) before its elementwise pass; both are pure HBM-bound
// reductions over rows.  These kernels compute exactly the quantities of ATen's batch_norm_stats /
// batch_norm_backward_reduce (the elementwise halves stay ATen's batch_norm_elemt / batch_norm_backward_elemt, which
// already run near the HBM rate), at about four times the speed of ATen's channels-last reduction kernels.
//
// Layout of a reduction: a thread owns VEC consecutive channels (one 16-byte load per row); the block is LX column
// threads x LY row lanes and walks a contiguous row range with LY rows per pass, several independent loads in flight
// per thread; row lanes are combined through shared memory and every block writes ONE partial per channel
// (no atomics: the result is deterministic).  A second, tiny kernel combines the partials.
//
// Numerics of the statistics: a block accumulates sums of d = x - pivot_c (pivot = the block's first row), so that
// sum d^2 - (sum d)^2 / n does not cancel when |mean| >> std; block partials (n, mean, M2) are merged in the finalize
// kernel as mean = sum n_b mean_b / N, M2 = sum [M2_b + n_b (mean_b - mean)^2] (Chan's parallel variance, all at once).
#include "knn_tc.cuh"

namespace gkg {
namespace {

constexpr int kBnMaxThreads = 512;
constexpr int kBnMaxBlocks = 148 * 2;      // row ranges (partials per channel)
constexpr int kBnUnroll = 4;               // independent row loads in flight per thread

template <typename T> struct BnVec;
template <> struct BnVec<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct BnVec<float> { static constexpr int N = 4; };

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) BnPack {
  T v[VEC];
};

// partial layout: [block][2][C] floats
// MODE 0: statistics (shifted sum, sum of squares); 1: backward reduce (sum dy, sum dy * (x - mean)); 2: column sum of x
// nn.GELU() (erf form) and its derivative from one evaluation of the Abramowitz-Stegun 7.1.26 polynomial (the
// arrangement of grouped_fc.cu: Phi(-|z|) = (0.5 poly(t)) 2^(-(|z| sqrt(log2(e)/2))^2), |error of Phi| <= 7.5e-8):
//   gelu(z) = z Phi(z),   gelu'(z) = Phi(z) + z phi(z),   phi(z) = exp(-z^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ void gelu_parts(float z, float& phi_cdf, float& pdf) {
  const float a = fabsf(z);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.231641888f, a, 1.f)));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float sq = a * 0.849321800f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-sq * sq));
  const float h = p * t * e;
  phi_cdf = z >= 0.f ? 1.f - h : h;
  pdf = e * 0.3989422804f;
}
__device__ __forceinline__ float gelu_fwd(float z) {
  float c, d;
  gelu_parts(z, c, d);
  return z * c;
}
__device__ __forceinline__ float gelu_grad(float z) {
  float c, d;
  gelu_parts(z, c, d);
  return fmaf(z, d, c);
}

// MODE 3: backward reduce THROUGH the activation: g = dy * gelu'(z), z = (x - mean) * (invstd * gamma) + beta
// (MODE 3 only: invstd, gamma, beta)
template <typename T, int MODE>
__global__ void __launch_bounds__(kBnMaxThreads)
bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                 float* __restrict__ partial, long long rows, int C, int LX, int LY, long long rows_per_block,
                 const float* __restrict__ invstd = nullptr, const float* __restrict__ gamma = nullptr,
                 const float* __restrict__ beta = nullptr, const T* __restrict__ dact = nullptr) {
  constexpr int VEC = BnVec<T>::N;
  constexpr bool BWD = MODE == 1 || MODE == 3 || MODE == 4, SUM = MODE == 2, ACTG = MODE == 3, SAVED = MODE == 4;
  using P = BnPack<T, VEC>;
  extern __shared__ float bn_s[];            // [LY][LX][2 * VEC]
  const int tx = threadIdx.x % LX, ty = threadIdx.x / LX;
  const int cchunk = blockIdx.y * LX + tx;   // 16-byte column chunk of this thread
  const int c0 = cchunk * VEC;
  const bool col_ok = c0 < C;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float a[VEC], b[VEC], piv[VEC], asc[ACTG ? VEC : 1], ash[ACTG ? VEC : 1];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { a[e] = 0.f; b[e] = 0.f; piv[e] = 0.f; }
  if (ACTG && col_ok) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) { asc[ACTG ? e : 0] = invstd[c0 + e] * gamma[c0 + e]; ash[ACTG ? e : 0] = beta[c0 + e]; }
  }
  if (col_ok && r0 < r1) {
    if (BWD) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) piv[e] = mean[c0 + e];
    } else if (!SUM) {
      const P p = *reinterpret_cast<const P*>(x + r0 * C + c0);
#pragma unroll
      for (int e = 0; e < VEC; ++e) piv[e] = to_f32<T>(p.v[e]);
    }
    long long r = r0 + ty;
    for (; r + (long long)(kBnUnroll - 1) * LY < r1; r += (long long)kBnUnroll * LY) {
      P xv[kBnUnroll], gv[kBnUnroll], av[SAVED ? kBnUnroll : 1];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        xv[u] = *reinterpret_cast<const P*>(x + (r + (long long)u * LY) * C + c0);
        if (BWD) gv[u] = *reinterpret_cast<const P*>(dy + (r + (long long)u * LY) * C + c0);
        if (SAVED) av[SAVED ? u : 0] = *reinterpret_cast<const P*>(dact + (r + (long long)u * LY) * C + c0);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float d = to_f32<T>(xv[u].v[e]) - piv[e];
          if (BWD) {
            float g = to_f32<T>(gv[u].v[e]);
            if (ACTG) g *= gelu_grad(fmaf(d, asc[ACTG ? e : 0], ash[ACTG ? e : 0]));
            if (SAVED) g *= to_f32<T>(av[SAVED ? u : 0].v[e]);
            a[e] += g;
            b[e] = fmaf(g, d, b[e]);
          } else {
            a[e] += d;
            if (!SUM) b[e] = fmaf(d, d, b[e]);
          }
        }
      }
    }
    for (; r < r1; r += LY) {
      const P xv = *reinterpret_cast<const P*>(x + r * C + c0);
      P gv, av;
      if (BWD) gv = *reinterpret_cast<const P*>(dy + r * C + c0);
      if (SAVED) av = *reinterpret_cast<const P*>(dact + r * C + c0);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float d = to_f32<T>(xv.v[e]) - piv[e];
        if (BWD) {
          float g = to_f32<T>(gv.v[e]);
          if (ACTG) g *= gelu_grad(fmaf(d, asc[ACTG ? e : 0], ash[ACTG ? e : 0]));
          if (SAVED) g *= to_f32<T>(av.v[e]);
          a[e] += g;
          b[e] = fmaf(g, d, b[e]);
        } else {
          a[e] += d;
          if (!SUM) b[e] = fmaf(d, d, b[e]);
        }
      }
    }
  }
  float* mine = bn_s + ((size_t)ty * LX + tx) * (2 * VEC);
#pragma unroll
  for (int e = 0; e < VEC; ++e) { mine[e] = a[e]; mine[VEC + e] = b[e]; }
  __syncthreads();
  if (ty == 0 && col_ok) {
    for (int l = 1; l < LY; ++l) {
      const float* o = bn_s + ((size_t)l * LX + tx) * (2 * VEC);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { a[e] += o[e]; b[e] += o[VEC + e]; }
    }
    float* out = partial + (size_t)blockIdx.x * 2 * C;
    const float n = (float)(r1 > r0 ? r1 - r0 : 0);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (BWD || SUM) {
        out[c0 + e] = a[e];                  // sum dy | column sum
        out[C + c0 + e] = b[e];              // sum dy * (x - mean)
      } else {
        // block mean and M2 = sum (x - block mean)^2 from the shifted sums
        const float md = n > 0.f ? a[e] / n : 0.f;
        out[c0 + e] = piv[e] + md;
        out[C + c0 + e] = fmaxf(b[e] - a[e] * md, 0.f);
      }
    }
  }
}

// Elementwise halves for the norm -> GELU pairs (BasicConv torch_nn.py:61-65, FFN.fc1 gkgnet.py:52-58, Stem): the
// activation is applied in the same pass as the normalisation, and undone (g = dy * gelu'(z), z recomputed from x) in
// the two backward passes -- the stand-alone GELU forward / backward passes and the saved norm output disappear.
// Same thread layout as the reductions: a thread keeps its 8 channels' coefficients in registers.
// FWD:  y = gelu((x - mean) * sc + beta),  sc = invstd * gamma
// BWD:  dx = (g - sum_g / N - (x - mean) * invstd^2 * sum_g_xmu / N) * sc
// SAVED: the forward also writes gelu'(z) (dact, feature dtype) and the backward passes multiply by it instead of
// re-evaluating the derivative -- with it the two backward passes are pure streaming kernels (no MUFU work at all).
template <typename T, bool BWD, bool SAVED>
__global__ void __launch_bounds__(kBnMaxThreads)
bn_act_elemt_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ out, T* __restrict__ dact,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ sum_g, const float* __restrict__ sum_g_xmu,
                    long long rows, int C, int LX, int LY, long long rows_per_block, float inv_n) {
  constexpr int VEC = BnVec<T>::N;
  using P = BnPack<T, VEC>;
  const int tx = threadIdx.x % LX, ty = threadIdx.x / LX;
  const int c0 = (blockIdx.y * LX + tx) * VEC;
  if (c0 >= C) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float mu[VEC], sc[VEC], sh[VEC], k1[BWD ? VEC : 1], k2[BWD ? VEC : 1];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const float is = invstd[c0 + e];
    mu[e] = mean[c0 + e];
    sc[e] = is * gamma[c0 + e];
    sh[e] = beta[c0 + e];
    if (BWD) {
      k1[BWD ? e : 0] = sum_g[c0 + e] * inv_n;
      k2[BWD ? e : 0] = is * is * sum_g_xmu[c0 + e] * inv_n;
    }
  }
  long long r = r0 + ty;
  for (; r + (long long)(kBnUnroll - 1) * LY < r1; r += (long long)kBnUnroll * LY) {
    P xv[kBnUnroll], gv[kBnUnroll], av[(BWD && SAVED) ? kBnUnroll : 1];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      xv[u] = *reinterpret_cast<const P*>(x + (r + (long long)u * LY) * C + c0);
      if (BWD) gv[u] = *reinterpret_cast<const P*>(dy + (r + (long long)u * LY) * C + c0);
      if (BWD && SAVED) av[(BWD && SAVED) ? u : 0] = *reinterpret_cast<const P*>(dact + (r + (long long)u * LY) * C + c0);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      P o, da;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float d = to_f32<T>(xv[u].v[e]) - mu[e];
        const float z = fmaf(d, sc[e], sh[e]);
        if (BWD) {
          const float g = to_f32<T>(gv[u].v[e]) * (SAVED ? to_f32<T>(av[(BWD && SAVED) ? u : 0].v[e]) : gelu_grad(z));
          o.v[e] = from_f32<T>((g - k1[BWD ? e : 0] - d * k2[BWD ? e : 0]) * sc[e]);
        } else {
          float cdf, pdf;
          gelu_parts(z, cdf, pdf);
          o.v[e] = from_f32<T>(z * cdf);
          if (SAVED) da.v[e] = from_f32<T>(fmaf(z, pdf, cdf));
        }
      }
      *reinterpret_cast<P*>(out + (r + (long long)u * LY) * C + c0) = o;
      if (!BWD && SAVED) *reinterpret_cast<P*>(dact + (r + (long long)u * LY) * C + c0) = da;
    }
  }
  for (; r < r1; r += LY) {
    const P xv = *reinterpret_cast<const P*>(x + r * C + c0);
    P gv, av, o, da;
    if (BWD) gv = *reinterpret_cast<const P*>(dy + r * C + c0);
    if (BWD && SAVED) av = *reinterpret_cast<const P*>(dact + r * C + c0);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float d = to_f32<T>(xv.v[e]) - mu[e];
      const float z = fmaf(d, sc[e], sh[e]);
      if (BWD) {
        const float g = to_f32<T>(gv.v[e]) * (SAVED ? to_f32<T>(av.v[e]) : gelu_grad(z));
        o.v[e] = from_f32<T>((g - k1[BWD ? e : 0] - d * k2[BWD ? e : 0]) * sc[e]);
      } else {
        float cdf, pdf;
        gelu_parts(z, cdf, pdf);
        o.v[e] = from_f32<T>(z * cdf);
        if (SAVED) da.v[e] = from_f32<T>(fmaf(z, pdf, cdf));
      }
    }
    *reinterpret_cast<P*>(out + r * C + c0) = o;
    if (!BWD && SAVED) *reinterpret_cast<P*>(dact + r * C + c0) = da;
  }
}

// Combine the block partials.  A block takes 32 channels x kFinLanes partial lanes: lane py merges partials py,
// py + kFinLanes, ... of its channel (reads coalesced over the 32 channels; a one-thread-per-channel loop over 592
// partials was a 100 us dependent chain of L2 loads), the lanes are merged through shared memory.
constexpr int kFinLanes = 16;

// Merge of the block partials (n_b, mean_b, M2_b) -> mean, invstd; running statistics like nn.BatchNorm2d (momentum
// update, unbiased variance).  Two passes over the (L2-resident) partials instead of a chain of Chan updates with two
// divisions each:  mean = sum n_b mean_b / N,  M2 = sum [M2_b + n_b (mean_b - mean)^2]  -- the same quantity, every term
// non-negative (no cancellation), one division per channel.
__global__ void __launch_bounds__(32 * kFinLanes)
bn_stats_finalize_kernel(const float* __restrict__ partial, int nblocks, long long rows, long long rows_per_block,
                         int C, float eps, float momentum, float* __restrict__ mean_out,
                         float* __restrict__ invstd_out, float* __restrict__ running_mean,
                         float* __restrict__ running_var) {
  __shared__ float sm[kFinLanes][32];
  __shared__ float s_mean[32];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < C;
  const float n_full = (float)rows_per_block;
  const float n_last = (float)(rows - (long long)(nblocks - 1) * rows_per_block);     // the last range may be shorter
  const float inv_n = 1.f / (float)rows;
  // pass 1: weighted mean, relative to the first block's mean (keeps the sum small when |mean| >> std)
  const float pivot = ok ? partial[c] : 0.f;
  float acc = 0.f;
  if (ok)
    for (int bI = py; bI < nblocks; bI += kFinLanes)
      acc = fmaf(bI == nblocks - 1 ? n_last : n_full, partial[(size_t)bI * 2 * C + c] - pivot, acc);
  sm[py][cx] = acc;
  __syncthreads();
  if (py == 0) {
    for (int l = 1; l < kFinLanes; ++l) acc += sm[l][cx];
    s_mean[cx] = pivot + acc * inv_n;
  }
  __syncthreads();
  const float mean = s_mean[cx];
  // pass 2: within-block + between-block sums of squares
  float m2 = 0.f;
  if (ok)
    for (int bI = py; bI < nblocks; bI += kFinLanes) {
      const float dm = partial[(size_t)bI * 2 * C + c] - mean;
      m2 += fmaf(bI == nblocks - 1 ? n_last : n_full, dm * dm, partial[(size_t)bI * 2 * C + C + c]);
    }
  __syncthreads();
  sm[py][cx] = m2;
  __syncthreads();
  if (py != 0 || !ok) return;
  for (int l = 1; l < kFinLanes; ++l) m2 += sm[l][cx];
  const float n = (float)rows;
  const float var = m2 * inv_n;
  mean_out[c] = mean;
  invstd_out[c] = rsqrtf(var + eps);
  if (running_mean != nullptr) {
    const float unbiased = n > 1.f ? m2 / (n - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

__global__ void __launch_bounds__(32 * kFinLanes)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, const float* __restrict__ invstd,
                       float* __restrict__ sum_dy, float* __restrict__ sum_dy_xmu, float* __restrict__ grad_weight,
                       float* __restrict__ grad_bias) {
  __shared__ float ss[kFinLanes][32], st[kFinLanes][32];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f, t = 0.f;
  if (c < C) {
    for (int bI = py; bI < nblocks; bI += kFinLanes) {
      s += partial[(size_t)bI * 2 * C + c];
      t += partial[(size_t)bI * 2 * C + C + c];
    }
  }
  ss[py][cx] = s; st[py][cx] = t;
  __syncthreads();
  if (py != 0 || c >= C) return;
  for (int l = 1; l < kFinLanes; ++l) { s += ss[l][cx]; t += st[l][cx]; }
  sum_dy[c] = s;
  if (sum_dy_xmu != nullptr) sum_dy_xmu[c] = t;
  if (grad_weight != nullptr) grad_weight[c] = t * invstd[c];
  if (grad_bias != nullptr) grad_bias[c] = s;
}

struct BnPlan {
  int LX, LY, slabs, blocks;
  long long rows_per_block;
  size_t smem;
};

BnPlan bn_plan(long long rows, int C, int vec) {
  BnPlan p;
  const int chunks = C / vec;
  p.slabs = (chunks + 255) / 256;            // column slabs of <= 256 chunks, evenly filled
  p.LX = (chunks + p.slabs - 1) / p.slabs;
  p.LY = kBnMaxThreads / p.LX;
  if (p.LY > 64) p.LY = 64;
  // enough row ranges to fill the GPU, not so many that a range is shorter than a few passes
  long long want = (rows + (long long)p.LY * kBnUnroll * 2 - 1) / ((long long)p.LY * kBnUnroll * 2);
  if (want < 1) want = 1;
  long long cap = kBnMaxBlocks / p.slabs;
  if (cap < 1) cap = 1;
  p.blocks = (int)(want < cap ? want : cap);
  p.rows_per_block = (rows + p.blocks - 1) / p.blocks;
  p.blocks = (int)((rows + p.rows_per_block - 1) / p.rows_per_block);
  p.smem = sizeof(float) * (size_t)p.LX * p.LY * 2 * vec;
  return p;
}

template <typename T, int MODE>
int launch_bn_reduce(const void* x, const void* dy, const float* mean, float* partial, long long rows, int C,
                     const BnPlan& p, cudaStream_t stream, const float* invstd = nullptr, const float* gamma = nullptr,
                     const float* beta = nullptr, const void* dact = nullptr) {
  static std::atomic<uint64_t> configured{0};
  configure_once_per_device(configured, [] {
    cudaFuncSetAttribute(bn_reduce_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  });
  dim3 grid(p.blocks, p.slabs);
  bn_reduce_kernel<T, MODE><<<grid, p.LX * p.LY, p.smem, stream>>>(static_cast<const T*>(x), static_cast<const T*>(dy),
                                                                 mean, partial, rows, C, p.LX, p.LY, p.rows_per_block,
                                                                 invstd, gamma, beta, static_cast<const T*>(dact));
  GKG_CHECK_LAUNCH(MODE == 4 ? "bn_reduce_kernel<bwd, saved gelu'>" : MODE == 3 ? "bn_reduce_kernel<bwd, gelu>" : MODE == 1 ? "bn_reduce_kernel<bwd>" : MODE == 2 ? "bn_reduce_kernel<colsum>" : "bn_reduce_kernel<stats>");
  return GKG_OK;
}

int check_bn_args(const void* x, long long rows, int C, int dtype, int* vec) {
  GKG_CHECK_ARG(dtype == GKG_F32 || dtype == GKG_BF16, "batch_norm: dtype %d", dtype);
  *vec = dtype == GKG_F32 ? 4 : 8;
  GKG_CHECK_ARG(rows >= 1 && C >= *vec && C % *vec == 0, "batch_norm: rows=%lld C=%d (C must be a multiple of %d)", rows, C,
                *vec);
  GKG_CHECK_ARG(((uintptr_t)x % 16) == 0, "batch_norm: activation pointer not 16-byte aligned");
  return GKG_OK;
}

}  // namespace
}  // namespace gkg

using namespace gkg;

extern "C" size_t gkg_bn_workspace_bytes(long long rows, int C) {
  (void)rows;                                           // per-block partials + the two sums of the fused backward
  return sizeof(float) * ((size_t)kBnMaxBlocks + 1) * 2 * (size_t)(C > 0 ? C : 0);
}

extern "C" int gkg_bn_stats(const void* x, long long rows, int C, int dtype, float eps, float momentum,
                            float* mean, float* invstd, float* running_mean, float* running_var, void* ws,
                            size_t ws_bytes, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int vec = 0;
  int rc = check_bn_args(x, rows, C, dtype, &vec);
  if (rc != GKG_OK) return rc;
  GKG_CHECK_ARG(mean != nullptr && invstd != nullptr && ws != nullptr, "gkg_bn_stats: null output / workspace");
  GKG_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "gkg_bn_stats: running_mean / running_var");
  GKG_CHECK_ARG(ws_bytes >= gkg_bn_workspace_bytes(rows, C), "gkg_bn_stats: workspace too small");
  const BnPlan p = bn_plan(rows, C, vec);
  float* partial = static_cast<float*>(ws);
  rc = dtype == GKG_F32 ? launch_bn_reduce<float, 0>(x, nullptr, nullptr, partial, rows, C, p, stream)
                        : launch_bn_reduce<__nv_bfloat16, 0>(x, nullptr, nullptr, partial, rows, C, p, stream);
  if (rc != GKG_OK) return rc;
  bn_stats_finalize_kernel<<<(C + 31) / 32, 32 * kFinLanes, 0, stream>>>(partial, p.blocks, rows, p.rows_per_block, C, eps,
                                                                momentum, mean, invstd, running_mean, running_var);
  GKG_CHECK_LAUNCH("bn_stats_finalize_kernel");
  return GKG_OK;
}

extern "C" int gkg_bn_backward_reduce(const void* grad_out, const void* x, const float* mean, const float* invstd,
                                      long long rows, int C, int dtype, float* sum_dy, float* sum_dy_xmu,
                                      float* grad_weight, float* grad_bias, void* ws, size_t ws_bytes,
                                      gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int vec = 0;
  int rc = check_bn_args(x, rows, C, dtype, &vec);
  if (rc != GKG_OK) return rc;
  GKG_CHECK_ARG(((uintptr_t)grad_out % 16) == 0, "gkg_bn_backward_reduce: gradient pointer not 16-byte aligned");
  GKG_CHECK_ARG(mean != nullptr && invstd != nullptr && sum_dy != nullptr && sum_dy_xmu != nullptr && ws != nullptr,
                "gkg_bn_backward_reduce: null pointer");
  GKG_CHECK_ARG(ws_bytes >= gkg_bn_workspace_bytes(rows, C), "gkg_bn_backward_reduce: workspace too small");
  const BnPlan p = bn_plan(rows, C, vec);
  float* partial = static_cast<float*>(ws);
  rc = dtype == GKG_F32 ? launch_bn_reduce<float, 1>(x, grad_out, mean, partial, rows, C, p, stream)
                        : launch_bn_reduce<__nv_bfloat16, 1>(x, grad_out, mean, partial, rows, C, p, stream);
  if (rc != GKG_OK) return rc;
  bn_bwd_finalize_kernel<<<(C + 31) / 32, 32 * kFinLanes, 0, stream>>>(partial, p.blocks, C, invstd, sum_dy, sum_dy_xmu,
                                                              grad_weight, grad_bias);
  GKG_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  return GKG_OK;
}

// Column sums of a (rows, C) activation: the bias gradient of a token-major 1x1 convolution / linear layer
// (what autograd derives for Conv2d(.., 1).bias in Grapher.fc1 / fc2, FFN: torch_vertex.py:290-306, gkgnet.py:46-72).
extern "C" int gkg_column_sum(const void* x, long long rows, int C, int dtype, float* out, void* ws, size_t ws_bytes,
                              gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int vec = 0;
  int rc = check_bn_args(x, rows, C, dtype, &vec);
  if (rc != GKG_OK) return rc;
  GKG_CHECK_ARG(out != nullptr && ws != nullptr, "gkg_column_sum: null output / workspace");
  GKG_CHECK_ARG(ws_bytes >= gkg_bn_workspace_bytes(rows, C), "gkg_column_sum: workspace too small");
  const BnPlan p = bn_plan(rows, C, vec);
  float* partial = static_cast<float*>(ws);
  rc = dtype == GKG_F32 ? launch_bn_reduce<float, 2>(x, nullptr, nullptr, partial, rows, C, p, stream)
                        : launch_bn_reduce<__nv_bfloat16, 2>(x, nullptr, nullptr, partial, rows, C, p, stream);
  if (rc != GKG_OK) return rc;
  bn_bwd_finalize_kernel<<<(C + 31) / 32, 32 * kFinLanes, 0, stream>>>(partial, p.blocks, C, nullptr, out, nullptr,
                                                                      nullptr, nullptr);
  GKG_CHECK_LAUNCH("bn_bwd_finalize_kernel<colsum>");
  return GKG_OK;
}

// norm -> GELU in one pass and its backward (see bn_act_elemt_kernel); act: 2 = GELU (erf form) -- the only pairing the
// reference's stacks contain.
extern "C" int gkg_bn_act_forward(const void* x, const float* mean, const float* invstd, const float* gamma,
                                  const float* beta, long long rows, int C, int dtype, int act, void* y, void* dact,
                                  gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int vec = 0;
  int rc = check_bn_args(x, rows, C, dtype, &vec);
  if (rc != GKG_OK) return rc;
  GKG_CHECK_ARG(act == 2, "gkg_bn_act_forward: activation %d (only 2 = GELU)", act);
  GKG_CHECK_ARG(mean && invstd && gamma && beta && y && ((uintptr_t)y % 16) == 0 && ((uintptr_t)dact % 16) == 0,
                "gkg_bn_act_forward: null / unaligned pointer");
  const BnPlan p = bn_plan(rows, C, vec);
  dim3 grid(p.blocks, p.slabs);
#define GKG_BN_FWD(TT, SV)                                                                                         \
  bn_act_elemt_kernel<TT, false, SV><<<grid, p.LX * p.LY, 0, stream>>>(                                            \
      static_cast<const TT*>(x), nullptr, static_cast<TT*>(y), static_cast<TT*>(dact), mean, invstd, gamma, beta,  \
      nullptr, nullptr, rows, C, p.LX, p.LY, p.rows_per_block, 0.f)
  if (dtype == GKG_F32) { if (dact) GKG_BN_FWD(float, true); else GKG_BN_FWD(float, false); }
  else { if (dact) GKG_BN_FWD(__nv_bfloat16, true); else GKG_BN_FWD(__nv_bfloat16, false); }
#undef GKG_BN_FWD
  GKG_CHECK_LAUNCH("bn_act_elemt_kernel<fwd>");
  return GKG_OK;
}

// The fused backward in its two halves (a data-parallel SyncBN all-reduces `sums` between them):
//   reduce: sums[0..C) = sum g, sums[C..2C) = sum g (x - mean), g = dy * gelu'(z); grad_weight / grad_bias from the
//           LOCAL sums (they are parameter gradients: the gradient all-reduce averages them like every other one)
//   elemt:  dx from the (possibly all-reduced) sums and the total row count they cover
extern "C" int gkg_bn_act_backward_reduce(const void* grad_out, const void* x, const void* dact, const float* mean,
                                          const float* invstd, const float* gamma, const float* beta, long long rows,
                                          int C, int dtype, int act, float* sums, float* grad_weight, float* grad_bias,
                                          void* ws, size_t ws_bytes, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int vec = 0;
  int rc = check_bn_args(x, rows, C, dtype, &vec);
  if (rc != GKG_OK) return rc;
  GKG_CHECK_ARG(act == 2, "gkg_bn_act_backward_reduce: activation %d (only 2 = GELU)", act);
  GKG_CHECK_ARG(grad_out && mean && invstd && gamma && beta && sums && ws, "gkg_bn_act_backward_reduce: null pointer");
  GKG_CHECK_ARG(((uintptr_t)grad_out % 16) == 0 && ((uintptr_t)dact % 16) == 0, "gkg_bn_act_backward_reduce: unaligned pointer");
  GKG_CHECK_ARG(ws_bytes >= gkg_bn_workspace_bytes(rows, C), "gkg_bn_act_backward_reduce: workspace too small");
  const BnPlan p = bn_plan(rows, C, vec);
  float* partial = static_cast<float*>(ws);
  if (dact != nullptr)
    rc = dtype == GKG_F32
             ? launch_bn_reduce<float, 4>(x, grad_out, mean, partial, rows, C, p, stream, invstd, gamma, beta, dact)
             : launch_bn_reduce<__nv_bfloat16, 4>(x, grad_out, mean, partial, rows, C, p, stream, invstd, gamma, beta, dact);
  else
    rc = dtype == GKG_F32
             ? launch_bn_reduce<float, 3>(x, grad_out, mean, partial, rows, C, p, stream, invstd, gamma, beta)
             : launch_bn_reduce<__nv_bfloat16, 3>(x, grad_out, mean, partial, rows, C, p, stream, invstd, gamma, beta);
  if (rc != GKG_OK) return rc;
  bn_bwd_finalize_kernel<<<(C + 31) / 32, 32 * kFinLanes, 0, stream>>>(partial, p.blocks, C, invstd, sums, sums + C,
                                                                      grad_weight, grad_bias);
  GKG_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  return GKG_OK;
}

extern "C" int gkg_bn_act_backward_elemt(const void* grad_out, const void* x, const void* dact, const float* mean,
                                         const float* invstd, const float* gamma, const float* beta, const float* sums,
                                         long long rows, long long total_rows, int C, int dtype, int act, void* grad_x,
                                         gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int vec = 0;
  int rc = check_bn_args(x, rows, C, dtype, &vec);
  if (rc != GKG_OK) return rc;
  GKG_CHECK_ARG(act == 2 && total_rows >= rows, "gkg_bn_act_backward_elemt: activation %d / total rows %lld", act, total_rows);
  GKG_CHECK_ARG(grad_out && mean && invstd && gamma && beta && sums && grad_x, "gkg_bn_act_backward_elemt: null pointer");
  GKG_CHECK_ARG(((uintptr_t)grad_out % 16) == 0 && ((uintptr_t)grad_x % 16) == 0 && ((uintptr_t)dact % 16) == 0,
                "gkg_bn_act_backward_elemt: unaligned pointer");
  const BnPlan p = bn_plan(rows, C, vec);
  const float inv_n = 1.f / (float)total_rows;
  dim3 grid(p.blocks, p.slabs);
#define GKG_BN_BWD(TT, SV)                                                                                          \
  bn_act_elemt_kernel<TT, true, SV><<<grid, p.LX * p.LY, 0, stream>>>(                                              \
      static_cast<const TT*>(x), static_cast<const TT*>(grad_out), static_cast<TT*>(grad_x),                        \
      const_cast<TT*>(static_cast<const TT*>(dact)), mean, invstd, gamma, beta, sums, sums + C, rows, C, p.LX, p.LY,  \
      p.rows_per_block, inv_n)
  if (dtype == GKG_F32) { if (dact) GKG_BN_BWD(float, true); else GKG_BN_BWD(float, false); }
  else { if (dact) GKG_BN_BWD(__nv_bfloat16, true); else GKG_BN_BWD(__nv_bfloat16, false); }
#undef GKG_BN_BWD
  GKG_CHECK_LAUNCH("bn_act_elemt_kernel<bwd>");
  return GKG_OK;
}

extern "C" int gkg_bn_act_backward(const void* grad_out, const void* x, const void* dact, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, long long rows, int C,
                                   int dtype, int act, void* grad_x, float* grad_weight, float* grad_bias, void* ws,
                                   size_t ws_bytes, gkg_stream_t stream_) {
  GKG_CHECK_ARG(ws != nullptr && ws_bytes >= gkg_bn_workspace_bytes(rows, C), "gkg_bn_act_backward: workspace too small");
  float* sums = static_cast<float*>(ws) + (size_t)kBnMaxBlocks * 2 * (size_t)(C > 0 ? C : 0);   // [2][C] behind the partials
  int rc = gkg_bn_act_backward_reduce(grad_out, x, dact, mean, invstd, gamma, beta, rows, C, dtype, act, sums,
                                      grad_weight, grad_bias, ws, ws_bytes, stream_);
  if (rc != GKG_OK) return rc;
  return gkg_bn_act_backward_elemt(grad_out, x, dact, mean, invstd, gamma, beta, sums, rows, rows, C, dtype, act,
                                   grad_x, stream_);
}

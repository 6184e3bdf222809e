// Label-query head of GKGNet: per-label scores and the two multi-label losses, forward and backward.
//
// Reference: LabelQueryHead.get_score (mmcls/models/heads/label_query_head.py:49-57) computes fc1 on every label
// embedding -- a (B, n, n) product -- and keeps the diagonal with an eye mask, then adds fc2(gap):
//     score[b, i] = W1[i] . L[b, i] + b1[i] + W2[i] . gap[b] + b2[i]
// forward_train (:70-85, double_loss) adds AsymmetricLoss x 10 (losses/asymmetric_loss.py:9-72: gamma_pos 0,
// gamma_neg 2, clip 0.05, eps 1e-8) and the label-smoothed BCE-with-logits (label_smooth_loss.py:122-126, 168-175),
// both summed over all entries and divided by the batch.  Here: one row-dot kernel for the scores (no (B, n, n)
// intermediate), one kernel for both losses and their derivatives, one backward kernel per gradient.  fp32 throughout
// (the tensors are tiny: (B, 80, 640)).
#include "common.cuh"

namespace gkg {
namespace {

// one warp per (b, i)
__global__ void label_score_fwd_kernel(const float* __restrict__ L, const float* __restrict__ gap,
                                       const float* __restrict__ W1, const float* __restrict__ b1,
                                       const float* __restrict__ W2, const float* __restrict__ b2,
                                       float* __restrict__ score, int B, int n, int C) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * n) return;
  const int i = (int)(w % n);
  const long long b = w / n;
  const float* l = L + w * C;
  const float* g = gap + b * C;
  const float* w1 = W1 + (size_t)i * C;
  const float* w2 = W2 + (size_t)i * C;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    s1 = fmaf(l[c], w1[c], s1);
    s2 = fmaf(g[c], w2[c], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) score[w] = (s1 + b1[i]) + (s2 + b2[i]);
}

// dL[b, i, c] = ds[b, i] * W1[i, c]
__global__ void label_score_bwd_dl_kernel(const float* __restrict__ ds, const float* __restrict__ W1,
                                          float* __restrict__ dL, long long total, int n, int C) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % C);
  const long long bi = e / C;
  dL[e] = ds[bi] * W1[(size_t)(bi % n) * C + c];
}

// thread per (i, c): dW1[i, c] = sum_b ds[b, i] L[b, i, c], dW2[i, c] = sum_b ds[b, i] gap[b, c];
// threads with c == 0 also reduce the bias gradients
__global__ void label_score_bwd_w_kernel(const float* __restrict__ ds, const float* __restrict__ L,
                                         const float* __restrict__ gap, float* __restrict__ dW1,
                                         float* __restrict__ dW2, float* __restrict__ db1, float* __restrict__ db2,
                                         int B, int n, int C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * C) return;
  const int i = e / C, c = e - i * C;
  float a1 = 0.f, a2 = 0.f, ab = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = ds[(size_t)b * n + i];
    a1 = fmaf(d, L[((size_t)b * n + i) * C + c], a1);
    a2 = fmaf(d, gap[(size_t)b * C + c], a2);
    ab += d;
  }
  dW1[e] = a1;
  dW2[e] = a2;
  if (c == 0) { db1[i] = ab; db2[i] = ab; }
}

// thread per (b, c): dgap[b, c] = sum_i ds[b, i] W2[i, c]
__global__ void label_score_bwd_gap_kernel(const float* __restrict__ ds, const float* __restrict__ W2,
                                           float* __restrict__ dgap, int B, int n, int C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * C) return;
  const int b = e / C, c = e - b * C;
  float a = 0.f;
  for (int i = 0; i < n; ++i) a = fmaf(ds[(size_t)b * n + i], W2[(size_t)i * C + c], a);
  dgap[e] = a;
}

// Both losses (sums over all entries) and their derivatives with respect to the score, one block.
//   asl  = -log(max(pt, eps)) * (1 - pt)^(gp t + gn (1 - t)),  pt = min(1 - p + clip, 1) (1 - t) + p t,  p = sigmoid(s)
//   bce  = softplus(s) - ts * s,  ts = t (1 - 2 smooth) + smooth
__global__ void multilabel_loss_kernel(const float* __restrict__ score, const float* __restrict__ target,
                                       float* __restrict__ sums, float* __restrict__ d_asl, float* __restrict__ d_bce,
                                       long long total, float gp, float gn, float clip, float eps, float smooth) {
  __shared__ float red[2][32];
  float la = 0.f, lb = 0.f;
  for (long long e = threadIdx.x; e < total; e += blockDim.x) {
    const float s = score[e], t = target[e];
    const float p = 1.f / (1.f + expf(-s));
    const float dp = p * (1.f - p);
    const float a_raw = 1.f - p + clip;
    const bool clipped = !(clip > 0.f) ? false : a_raw > 1.f;
    const float a = clip > 0.f ? fminf(a_raw, 1.f) : 1.f - p;
    const float pt = a * (1.f - t) + p * t;
    const float dpt_dp = (clipped ? 0.f : -(1.f - t)) + t;
    const float gamma = gp * t + gn * (1.f - t);
    const float om = 1.f - pt;
    const float w = powf(om, gamma);
    const float ptc = fmaxf(pt, eps);
    const float lg = logf(ptc);
    la += -lg * w;
    // d/dpt [ -log(max(pt, eps)) (1 - pt)^gamma ]
    float dl = (pt > eps ? -w / pt : 0.f);
    if (gamma != 0.f) dl += lg * gamma * powf(om, gamma - 1.f);
    d_asl[e] = dl * dpt_dp * dp;
    const float ts = t * (1.f - 2.f * smooth) + smooth;
    lb += fmaxf(s, 0.f) - ts * s + log1pf(expf(-fabsf(s)));
    d_bce[e] = p - ts;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    la += __shfl_xor_sync(0xffffffffu, la, o);
    lb += __shfl_xor_sync(0xffffffffu, lb, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = la; red[1][threadIdx.x >> 5] = lb; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float a = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[threadIdx.x][w];
    sums[threadIdx.x] = a;
  }
}

}  // namespace
}  // namespace gkg

using namespace gkg;

extern "C" int gkg_label_score_fwd(const float* L, const float* gap, const float* W1, const float* b1, const float* W2,
                                   const float* b2, float* score, int B, int n, int C, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(B >= 0 && n > 0 && C > 0, "label_score_fwd: bad shape B=%d n=%d C=%d", B, n, C);
  if (B == 0) return GKG_OK;
  GKG_CHECK_ARG(L && gap && W1 && b1 && W2 && b2 && score, "label_score_fwd: null pointer");
  const long long warps = (long long)B * n;
  label_score_fwd_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(L, gap, W1, b1, W2, b2, score, B, n, C);
  GKG_CHECK_LAUNCH("label_score_fwd_kernel");
  return GKG_OK;
}

extern "C" int gkg_label_score_bwd(const float* dscore, const float* L, const float* gap, const float* W1,
                                   const float* W2, float* dL, float* dgap, float* dW1, float* dW2, float* db1,
                                   float* db2, int B, int n, int C, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(B >= 0 && n > 0 && C > 0, "label_score_bwd: bad shape B=%d n=%d C=%d", B, n, C);
  if (B == 0) return GKG_OK;
  GKG_CHECK_ARG(dscore && L && gap && W1 && W2 && dL && dgap && dW1 && dW2 && db1 && db2, "label_score_bwd: null pointer");
  const long long total = (long long)B * n * C;
  label_score_bwd_dl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dscore, W1, dL, total, n, C);
  GKG_CHECK_LAUNCH("label_score_bwd_dl_kernel");
  label_score_bwd_w_kernel<<<(n * C + 255) / 256, 256, 0, stream>>>(dscore, L, gap, dW1, dW2, db1, db2, B, n, C);
  GKG_CHECK_LAUNCH("label_score_bwd_w_kernel");
  label_score_bwd_gap_kernel<<<(B * C + 255) / 256, 256, 0, stream>>>(dscore, W2, dgap, B, n, C);
  GKG_CHECK_LAUNCH("label_score_bwd_gap_kernel");
  return GKG_OK;
}

extern "C" int gkg_multilabel_loss(const float* score, const float* target, float* sums, float* d_asl, float* d_bce,
                                   long long total, float gamma_pos, float gamma_neg, float clip, float eps,
                                   float smooth, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(total >= 0, "multilabel_loss: bad size %lld", total);
  GKG_CHECK_ARG(score && target && sums && d_asl && d_bce, "multilabel_loss: null pointer");
  multilabel_loss_kernel<<<1, 1024, 0, stream>>>(score, target, sums, d_asl, d_bce, total, gamma_pos, gamma_neg, clip, eps,
                                                 smooth);
  GKG_CHECK_LAUNCH("multilabel_loss_kernel");
  return GKG_OK;
}

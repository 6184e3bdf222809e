// Fused pairwise-distance + top-k on the Blackwell tensor cores (tcgen05 / TMEM / TMA bulk).
//
// What it computes (reference: torch_edge.py:39-51, 89-106, 139-149): for every query row n
// of problem p, the k*dilation keys with the smallest
//     dist[n, m] = (|xh_n|^2 - 2 xh_n . yh_m) + |yh_m|^2 + relative_pos[n, m]
// in ascending order, every `dilation`-th kept.  The N x M matrix never leaves the SM:
//
//   * operands: gkg_knn_prepare normalises the rows, splits them into fp16 hi/lo parts (scaled by 256 so the lo
//     part stays normal) and writes them K-concatenated,
//         A = [x_hi | S  S | 0.. | x_hi | x_lo],   B = [y_hi | c_hi c_lo | 0.. | y_lo | y_hi],   c = -0.5*|yh|^2*S,
//     so ONE fp16 GEMM with fp32 accumulation yields S^2 (xh.yh - |yh|^2/2) to ~2^-21 relative (the "fp16x3" split),
//     and its first PA columns alone the single-plane product the threshold sweep uses.  Rows are stored in the UMMA
//     no-swizzle K-major core-matrix order, tile by tile, so an operand block is one contiguous TMA bulk copy.  Wide
//     groups (D > 80, D % 8 == 0) store TWO planes per row instead, [hi (+ extras) | lo] for queries and keys alike --
//     the hi plane of the concatenated row is a repeat -- and the issuer forms hi.hi + lo.hi + hi.lo from them
//     (knn_tc_kernel.cuh: "split" plans);
//   * knn_tc_kernel (persistent, warp specialised, see knn_tc_kernel.cuh): warp 0 streams operand blocks, one warp per
//     row set (256-row items) or three warps taking the key tiles in turn (128-row items) issue tcgen05.mma into TMEM
//     accumulators, the epilogue warps drain them with tcgen05.ld (thread ==
//     query row), add the position bias and select in two sweeps over the keys of an item: a threshold sweep (sorted
//     register list of group maxima) and a logging sweep (predicated 16-byte shared-memory stores of the key triplets
//     that beat the threshold); the logged keys that reach the threshold go to global memory;
//   * knn_finalize_kernel sorts the handful of candidates per row, certifies the approximate order by its gaps
//     (>= 2 delta) and writes the dilated pick; rows that fail are re-ranked with the exact fp32 formula
//     (knn_rerank_kernel: same arithmetic as knn_exact.cu, the normalised query row rebuilt from the raw features),
//     rows whose candidate set itself is in doubt go to an exact brute-force kernel (knn_fixup_kernel).
#include "knn_tc_kernel.cuh"

namespace gkg {

using namespace tc;

namespace {

constexpr float kEps = 1e-12f;     // F.normalize eps

// ------------------------------------------------------------------------------------
// operand preparation (phase "prepare"): normalise + split + lay out, one pass over the features
// ------------------------------------------------------------------------------------
// A block owns 32 consecutive padded rows of one problem (a tile has tile_rows row slots of which the
// first tile_keys hold nodes: UMMA N must be a multiple of 16 while key tiles follow the bias period).
// Phase 1: a warp per row L2-normalises
// the D group channels (F.normalize semantics, torch_edge.py:167-168,173), keeps the fp32 row in
// shared memory and writes it + |xh|^2 for the exact re-rank.  Phase 2: every thread emits 16-byte
// core-matrix rows (8 fp16) of the tensor-core operand:
//   [tile][k-block][row group (tile_rows/8)][k chunk (KC/8)][row (8)][elem (8)]
// so that a warp writes 4 contiguous 128-byte core matrices.
constexpr int PREP_ROWS = 32;
// Row stride of the staged rows in shared memory (floats).  D % 8 == 0: phase 2 reads 8-column pieces with two
// LDS.128; the 8 lanes of a quarter warp read the same piece of 8 consecutive rows, so the stride in 16-byte units
// must be odd (D = 200: 50 -> 51).  Otherwise (scalar reads): stride == 1 (mod 8) floats spreads the 8 rows x 4
// pieces of a warp over all 32 banks.
__host__ __device__ constexpr int prep_stride(int D) {
  return D % 8 == 0 ? ((D / 4) % 2 == 0 ? D + 4 : D) : D + ((1 - D % 8) + 8) % 8;
}

// DT: group width at compile time (the wide stages: 200, 320 -- every index split below becomes a constant
// division), 0 = run-time value.
template <typename T, bool IS_KEY, int DT>
__global__ void __launch_bounds__(256)
tc_prepare_kernel(const T* __restrict__ feat, int64_t stride_b, int64_t stride_n, float* __restrict__ hat,
                  float* __restrict__ sq, __half* __restrict__ op, int G, int rows, int D_rt, int KP_rt, int KC,
                  int tiles, int tile_rows, int tile_keys, int write_hat, int PA_split) {
  // PA_split > 0: split-plane layout [hi (+ extra pair) -> PA_split | lo -> KP_rt] (knn_tc_kernel.cuh), same for
  // queries and keys; 0: the K-concatenated fp16x3 rows.
  const int D = DT > 0 ? DT : D_rt;
  const int KP = (DT > 0 && PA_split == 0) ? k_padded(DT > 0 ? DT : 1) : KP_rt;
  extern __shared__ __align__(16) float prep_s[];   // [PREP_ROWS][DS] normalised rows + [PREP_ROWS] norms
  const int DS = prep_stride(D);
  float* xs = prep_s;
  float* sqs = prep_s + PREP_ROWS * DS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long p = blockIdx.y;
  const int g = (int)(p % G);
  const long long b = p / G;
  const int r0 = blockIdx.x * PREP_ROWS;      // first PADDED row: tile_rows slots per tile, tile_keys of them real
  auto key_of = [&](int prow) {
    const int tile = prow / tile_rows, slot = prow - tile * tile_rows;
    return slot < tile_keys ? tile * tile_keys + slot : rows;   // rows == "no such node"
  };

  for (int rl = warp; rl < PREP_ROWS; rl += 8) {
    const int row = key_of(r0 + rl);
    float* dst = xs + rl * DS;
    if (row < rows) {
      const T* src = feat + b * stride_b + (long long)row * stride_n + (long long)g * D;
      float ss = 0.f;
      constexpr int NV = DT > 0 ? (DT + 31) / 32 : 1;     // compile-time width: the lane's elements stay in registers
      float xv[NV];
      if (DT > 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = lane + 32 * i;
          xv[i] = d < D ? to_f32<T>(src[d]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (lane + 32 * i < D) ss = fmaf(xv[i], xv[i], ss);
      } else {
        for (int d = lane; d < D; d += 32) {
          const float v = to_f32<T>(src[d]);
          ss = fmaf(v, v, ss);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float denom = fmaxf(sqrtf(ss), 1e-12f);
      float s2 = 0.f;
      float* gh = hat + (p * rows + row) * (long long)D;
      if (DT > 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = lane + 32 * i;
          if (d < D) {
            const float v = xv[i] / denom;
            dst[d] = v;
            if (write_hat) gh[d] = v;
            s2 = fmaf(v, v, s2);
          }
        }
      } else
      for (int d = lane; d < D; d += 32) {
        const float v = to_f32<T>(src[d]) / denom;
        dst[d] = v;
        if (write_hat) gh[d] = v;
        s2 = fmaf(v, v, s2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      if (lane == 0) {
        sqs[rl] = s2;
        if (write_hat) sq[p * rows + row] = s2;
      }
    } else {
      for (int d = lane; d < D; d += 32) dst[d] = 0.f;
      if (lane == 0) sqs[rl] = 0.f;
    }
  }
  __syncthreads();

  const int kcs = KC >> 3, nkb = KP / KC, rgs = tile_rows >> 3, kc_all = KP >> 3;
  const int PA = PA_split > 0 ? PA_split : k_prefix(D);
  if ((D & 7) == 0) {
    // One thread per (row, 8-column piece of the normalised row): two LDS.128, hi / lo split once, then the three
    // 16-byte core-matrix rows that piece feeds (A = [hi | hi | lo], B = [hi | lo | hi]); the remaining chunks of a row
    // (the extra column pair behind the first segment, zero padding) are a few more items.  Same arithmetic, hence the
    // same operand bits, as the per-chunk loop below -- at a third of its instructions.
    const int npiece = D >> 3;
    const int nmid = (PA - D) >> 3;                                 // chunks between the first segment and PA
    const int ntail0 = PA_split > 0 ? (PA + D) >> 3 : (PA + 2 * D) >> 3;   // first chunk of the zero tail
    const int nother = nmid + (kc_all - ntail0);
    const int per_row = npiece + nother;
    const int items = PREP_ROWS * per_row;
    for (int ci = threadIdx.x; ci < items; ci += 256) {
      const int r = ci & 7;
      const int it = (ci >> 3) % per_row;
      const int rl = ((ci >> 3) / per_row) * 8 + r;
      const int prow = r0 + rl;
      const int tile = prow / tile_rows;
      if (tile >= tiles) continue;
      const int rg = (prow - tile * tile_rows) >> 3;
      const long long tbase = (p * tiles + tile) * (long long)nkb;
      auto store = [&](int kcI, const uint4& val) {
        const int kb = kcI / kcs, kc = kcI - kb * kcs;
        const long long chunk = (((tbase + kb) * rgs + rg) * (long long)kcs + kc) * 8 + r;
        *reinterpret_cast<uint4*>(op + chunk * 8) = val;
      };
      if (it < npiece) {
        const float4 f0 = *reinterpret_cast<const float4*>(xs + rl * DS + it * 8);
        const float4 f1 = *reinterpret_cast<const float4*>(xs + rl * DS + it * 8 + 4);
        const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = f[e] * kScale;
          hi[e] = __float2half_rn(x);
          lo[e] = __float2half_rn(x - __half2float(hi[e]));
        }
        const uint4 H = *reinterpret_cast<const uint4*>(hi), Lo = *reinterpret_cast<const uint4*>(lo);
        store(it, H);
        if (PA_split > 0) {
          store((PA >> 3) + it, Lo);
        } else {
          store((PA >> 3) + it, IS_KEY ? Lo : H);
          store(((PA + D) >> 3) + it, IS_KEY ? H : Lo);
        }
      } else {
        const int o = it - npiece;
        const int kcI = o < nmid ? npiece + o : ntail0 + (o - nmid);
        __align__(16) __half z[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) z[e] = __float2half_rn(0.f);
        if (kcI == npiece) {                                          // columns D, D + 1: the extra pair
          const bool valid = key_of(prow) < rows;
          float extra_hi, extra_lo;
          if (IS_KEY) {
            if (valid) {
              const float c = -0.5f * sqs[rl] * kScale;
              extra_hi = __half2float(__float2half_rn(c));
              extra_lo = c - extra_hi;
            } else {
              extra_hi = kPadKey; extra_lo = 0.f;
            }
          } else {
            extra_hi = valid ? kScale : 0.f;
            extra_lo = extra_hi;
          }
          z[0] = __float2half_rn(extra_hi);
          z[1] = __float2half_rn(extra_lo);
        }
        store(kcI, *reinterpret_cast<const uint4*>(z));
      }
    }
    return;
  }
  const int total = PREP_ROWS * kc_all;
  for (int ci = threadIdx.x; ci < total; ci += 256) {
    const int r = ci & 7;
    const int kcI = (ci >> 3) % kc_all;
    const int rgl = (ci >> 3) / kc_all;
    const int rl = rgl * 8 + r;
    const int prow = r0 + rl;
    const int tile = prow / tile_rows;
    if (tile >= tiles) continue;
    const int rg = (prow - tile * tile_rows) >> 3;
    const int kb = kcI / kcs, kc = kcI - kb * kcs;
    const bool valid = key_of(prow) < rows;
    const float* src = xs + rl * DS;
    float extra_hi = 0.f, extra_lo = 0.f;
    if (IS_KEY) {
      if (valid) {
        const float c = -0.5f * sqs[rl] * kScale;
        extra_hi = __half2float(__float2half_rn(c));
        extra_lo = c - extra_hi;
      } else {
        extra_hi = kPadKey;
      }
    } else {
      extra_hi = valid ? kScale : 0.f;
      extra_lo = extra_hi;
    }
    __align__(16) __half out[8];
    const int c0 = kcI * 8;
    // columns: [hi (D) | e_hi e_lo | 0.. -> PA | second (D) | third (D) | 0..];  A = [hi | hi | lo], B = [hi | lo | hi]
    if ((D & 7) == 0) {
      // 8-column pieces never straddle a segment: one decision per piece, no per-element compares
      const int seg = c0 < D ? 0 : (c0 >= PA && c0 < PA + D) ? 1 : (c0 >= PA + D && c0 < PA + 2 * D) ? 2 : -1;
      if (seg >= 0) {
        const float* s8 = src + (seg == 0 ? c0 : c0 - PA - (seg - 1) * D);
        const bool want_lo = IS_KEY ? (seg == 1) : (seg == 2);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = s8[e] * kScale;
          const __half h = __float2half_rn(x);
          out[e] = want_lo ? __float2half_rn(x - __half2float(h)) : h;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = __float2half_rn(0.f);
        if (c0 == D) { out[0] = __float2half_rn(extra_hi); out[1] = __float2half_rn(extra_lo); }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        float v = 0.f;
        const int seg = c < D ? 0 : (c >= PA && c < PA + D) ? 1 : (c >= PA + D && c < PA + 2 * D) ? 2 : -1;
        if (seg >= 0) {
          const float x = src[seg == 0 ? c : c - PA - (seg - 1) * D] * kScale;
          const float hi = __half2float(__float2half_rn(x));
          const bool want_lo = IS_KEY ? (seg == 1) : (seg == 2);
          v = want_lo ? (x - hi) : hi;
        } else if (c == D) {
          v = extra_hi;
        } else if (c == D + 1) {
          v = extra_lo;
        }
        out[e] = __float2half_rn(v);
      }
    }
    const long long chunk = ((((p * tiles + tile) * nkb + kb) * rgs + rg) * (long long)kcs + kc) * 8 + r;
    *reinterpret_cast<uint4*>(op + chunk * 8) = *reinterpret_cast<const uint4*>(out);
  }
}

// Row-per-thread variant for small compile-time D (the large layers: D = 20 / 40 / 80): the whole row
// lives in registers, loads and stores are 128-bit, and every K column's (segment, element) is known at
// compile time -- ~10x fewer instructions than the generic kernel above.  Bitwise identical results:
// the squared sums replay the lane-strided partial sums + xor-shuffle tree of the warp-per-row kernels.
template <int D>
__device__ __forceinline__ float sumsq_warp_order(const float (&v)[D]) {
  float part[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    part[l] = 0.f;
#pragma unroll
    for (int d = l; d < D; d += 32) part[l] = fmaf(v[d], v[d], part[l]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int l = 0; l < o; ++l) part[l] += part[l + o];
  }
  return part[0];
}

template <typename T, int D> struct RowLoad;
template <int D> struct RowLoad<__nv_bfloat16, D> {
  static constexpr int kAlignElems = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* src, float (&v)[D]) {
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + c);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[c * 8 + 2 * e] = __uint_as_float(w[e] << 16);
        v[c * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
      }
    }
  }
};
template <int D> struct RowLoad<float, D> {
  static constexpr int kAlignElems = 4;
  static __device__ __forceinline__ void load(const float* src, float (&v)[D]) {
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(src) + c);
      v[c * 4] = q.x; v[c * 4 + 1] = q.y; v[c * 4 + 2] = q.z; v[c * 4 + 3] = q.w;
    }
  }
};

constexpr int PREP_THREADS = 128;

template <typename T, int D, bool IS_KEY>
__global__ void __launch_bounds__(PREP_THREADS)
tc_prepare_rows_kernel(const T* __restrict__ feat, int64_t stride_b, int64_t stride_n, float* __restrict__ hat,
                       float* __restrict__ sq, __half* __restrict__ op, int G, int rows, int KC, int tiles,
                       int tile_rows, int tile_keys, int write_hat) {
  constexpr int KP = k_padded(D), PA = k_prefix(D);
  const long long p = blockIdx.y;
  const int g = (int)(p % G);
  const long long b = p / G;
  const int prow = blockIdx.x * PREP_THREADS + threadIdx.x;
  const int tile = prow / tile_rows, slot = prow - tile * tile_rows;
  if (tile >= tiles) return;
  const int row = slot < tile_keys ? tile * tile_keys + slot : rows;
  const bool valid = row < rows;

  float v[D];
  float s2 = 0.f;
  if (valid) {
    RowLoad<T, D>::load(feat + b * stride_b + (long long)row * stride_n + (long long)g * D, v);
    const float denom = fmaxf(sqrtf(sumsq_warp_order<D>(v)), 1e-12f);
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = v[d] / denom;     // true division, like aten::div
    s2 = sumsq_warp_order<D>(v);
    if (write_hat) {
      float4* gh = reinterpret_cast<float4*>(hat + (p * rows + row) * (long long)D);
#pragma unroll
      for (int c = 0; c < D / 4; ++c) gh[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      sq[p * rows + row] = s2;
    }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = 0.f;
  }
  float extra_hi, extra_lo;
  if (IS_KEY) {
    if (valid) {
      const float c = -0.5f * s2 * kScale;
      extra_hi = __half2float(__float2half_rn(c));
      extra_lo = c - extra_hi;
    } else {
      extra_hi = kPadKey;
      extra_lo = 0.f;
    }
  } else {
    extra_hi = valid ? kScale : 0.f;
    extra_lo = extra_hi;
  }
  const int kcs = KC >> 3, nkb = KP / KC, rgs = tile_rows >> 3;
  const int rg = slot >> 3, r = slot & 7;
  uint4* base = reinterpret_cast<uint4*>(op) + ((p * tiles + tile) * nkb * (long long)rgs) * kcs * 8;
  // queries: the block IS one operand tile (tile_rows == PREP_THREADS), contiguous in the operand buffer -> it is assembled in
  // shared memory and leaves as ONE bulk store instead of KP / 8 16-byte stores per thread
  extern __shared__ __align__(128) uint4 prep_stage[];
  uint4* dst = IS_KEY ? base : prep_stage;
#pragma unroll
  for (int kcI = 0; kcI < KP / 8; ++kcI) {
    __align__(16) __half out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = kcI * 8 + e;                  // compile-time after unrolling
      // columns: [hi (D) | e_hi e_lo | 0.. -> PA | second (D) | third (D) | 0..];  A = [hi | hi | lo], B = [hi | lo | hi]
      float val = 0.f;
      const int seg = c < D ? 0 : (c >= PA && c < PA + D) ? 1 : (c >= PA + D && c < PA + 2 * D) ? 2 : -1;
      if (seg >= 0) {
        const float x = v[seg == 0 ? c : c - PA - (seg - 1) * D] * kScale;
        const float hi = __half2float(__float2half_rn(x));
        const bool want_lo = IS_KEY ? (seg == 1) : (seg == 2);
        val = want_lo ? (x - hi) : hi;
      } else if (c == D) {
        val = extra_hi;
      } else if (c == D + 1) {
        val = extra_lo;
      }
      out[e] = __float2half_rn(val);
    }
    const int kb = kcI / kcs, kc = kcI - kb * kcs;
    dst[(((long long)kb * rgs + rg) * kcs + kc) * 8 + r] = *reinterpret_cast<const uint4*>(out);
  }
  if (!IS_KEY) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(base), "r"((uint32_t)__cvta_generic_to_shared(prep_stage)), "r"((uint32_t)(tile_rows * KP * 2)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

// ------------------------------------------------------------------------------------
// sorted candidate list (registers)
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// final selection: candidates of a row -> sorted neighbour ids (one thread per row)
// ------------------------------------------------------------------------------------
// knn_tc_kernel leaves, per row, the keys whose fp16x3 score reaches the row's threshold (>= k*d + 1 of
// them, ~T on average; every key that is not listed scores <= the threshold).  Here: the L = T + 3 best by
// branch-free sorted insertion, the gap test that certifies the approximate order (gaps >= 2 delta), the
// dilated pick.  Rows that fail the test go to knn_rerank_kernel, rows without a trustworthy candidate set
// to the brute-force fix-up.  Layout of `cand`: [item][slot][row of the item] so that loads coalesce.
template <int L>
__global__ void __launch_bounds__(256)
knn_finalize_kernel(const float2* __restrict__ cand, const int* __restrict__ cand_count,
                    const float* __restrict__ cand_thr, int cand_slots, int rows_per_item, int row_sets, int QI,
                    int32_t* __restrict__ idx_out, int* rr_count, int* rr_list, int rr_cap, int* fix_count,
                    int* fix_rows, unsigned int* stats, long long total_rows, int N, int M, int k, int dilation,
                    float delta, int force_rerank) {
  const long long gr = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // item * rows_per_item + row
  if (gr >= total_rows) return;
  const int item = (int)(gr / rows_per_item), r = (int)(gr - (long long)item * rows_per_item);
  const int p = item / QI, qi = item - p * QI;
  const int n = (qi * row_sets + r / BM) * BM + (r % BM);
  if (n >= N) return;
  const int row = p * N + n;
  const int np = cand_count[gr];
  if (force_rerank == 3 || np < 0 || (np < k * dilation + 1 && np < M)) {   // no trustworthy candidate set (3: test hook)
    if (force_rerank != 0) atomicAdd(stats + 0, 1u);
    fix_rows[atomicAdd(fix_count, 1)] = row;
    return;
  }
  float v[L];
  int id[L];
#pragma unroll
  for (int j = 0; j < L; ++j) { v[j] = -INFINITY; id[j] = 0x7fffffff; }
  float rej = cand_thr[gr];                     // best approximate score among the keys that are not in the list
  const float2* src = cand + (size_t)item * cand_slots * rows_per_item + r;
  // Short lists (k*d <= 18): all NS candidate slots are loaded at once (independent loads; slots past np read as
  // -inf) and sorted by Batcher's merge-exchange network -- 103 compare-exchanges for 20 slots, 186 for 31, five ALU
  // operations each -- instead of np insertions into a sorted list of L (6 operations per list slot and candidate:
  // the kernel was bound by the ALU pipe).  Equal scores may come out in either order: their gap is below 2*delta,
  // so the row is re-ranked exactly anyway.
  constexpr int NS = L == 14 ? 20 : (L == 23 ? 31 : 0);
  if (NS > 0 && cand_slots == NS) {
    constexpr int NSA = NS > 0 ? NS : 1;
    float sv[NSA];
    int sid[NSA];
#pragma unroll
    for (int s = 0; s < NSA; ++s) {
      float2 c = make_float2(-INFINITY, __int_as_float(0x7fffffff));
      if (s < np) c = src[(size_t)s * rows_per_item];
      sv[s] = c.x;
      sid[s] = __float_as_int(c.y);
    }
#pragma unroll
    for (int p2 = 1; p2 < NSA; p2 <<= 1) {
#pragma unroll
      for (int kk = p2; kk >= 1; kk >>= 1) {
#pragma unroll
        for (int j = kk % p2; j <= NSA - 1 - kk; j += 2 * kk) {
#pragma unroll
          for (int i = 0; i <= (kk - 1 < NSA - j - kk - 1 ? kk - 1 : NSA - j - kk - 1); ++i) {
            if ((i + j) / (2 * p2) == (i + j + kk) / (2 * p2)) {
              const int a = i + j, b = i + j + kk;
              const bool sw = sv[b] > sv[a];
              const float fa = sv[a], fb = sv[b];
              const int ia = sid[a], ib = sid[b];
              sv[a] = sw ? fb : fa; sv[b] = sw ? fa : fb;
              sid[a] = sw ? ib : ia; sid[b] = sw ? ia : ib;
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < L; ++j) { v[j] = sv[j < NSA ? j : 0]; id[j] = sid[j < NSA ? j : 0]; }
    if (L < NSA) rej = fmaxf(rej, sv[L < NSA ? L : 0]);   // the best of what did not make the list
  } else
  for (int s = 0; s < np; ++s) {
    const float2 c = src[(size_t)s * rows_per_item];
    const float x = c.x;
    const int xid = __float_as_int(c.y);
    rej = fmaxf(rej, fminf(x, v[L - 1]));       // what drops off the end (or fails to enter)
#pragma unroll
    for (int j = L - 1; j >= 1; --j) {
      const bool up = x > v[j - 1];             // the slot above moves down
      const bool in = x > v[j];
      id[j] = up ? id[j - 1] : (in ? xid : id[j]);
      v[j] = up ? v[j - 1] : (in ? x : v[j]);
    }
    if (x > v[0]) { v[0] = x; id[0] = xid; }
  }
  const int kd = k * dilation;
  bool amb = force_rerank > 0;
#pragma unroll
  for (int j = 0; j + 1 < L; ++j)
    if (j < kd && (v[j] - v[j + 1]) * (-kScoreToDist) < 2.f * delta) amb = true;
  if (amb) {
    atomicAdd(stats + 0, 1u);
    const int slot = atomicAdd(rr_count, 1);
    if (slot < rr_cap) {
      int* dst = rr_list + (size_t)slot * (2 * L + 3);
      dst[0] = row;
      dst[1] = np < L ? np : L;
      dst[2] = __float_as_int(rej * kScoreToDist);   // every key outside the list has approx dist >= this
#pragma unroll
      for (int j = 0; j < L; ++j) { dst[3 + j] = id[j]; dst[3 + L + j] = __float_as_int(v[j] * kScoreToDist); }
    } else {
      fix_rows[atomicAdd(fix_count, 1)] = row;
    }
  }
  int32_t* out = idx_out + (size_t)row * k;
#pragma unroll
  for (int j = 0; j < L; ++j)
    if (j < kd && j % dilation == 0) out[j / dilation] = id[j];
}

// ------------------------------------------------------------------------------------
// exact arithmetic of the slow paths: the normalised query row rebuilt from the raw features
// ------------------------------------------------------------------------------------
struct RawFeat {
  const void* x; int64_t sb, sn; int dtype; int G;
};
// A warp normalises group row (p, n) into dst[0..D) with the arithmetic of knn_prepare.cu (lane-strided sums,
// xor-shuffle tree, true division); returns |xh|^2.
__device__ __forceinline__ float normalise_row_warp(const RawFeat& f, int p, int n, int D, float* dst, int lane) {
  const int g = p % f.G;
  const long long b = p / f.G;
  const long long off = b * f.sb + (long long)n * f.sn + (long long)g * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = f.dtype == GKG_F32 ? static_cast<const float*>(f.x)[off + d]
                                       : __bfloat162float(static_cast<const __nv_bfloat16*>(f.x)[off + d]);
    dst[d] = v;
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), kEps);
  float s2 = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = dst[d] / denom;
    dst[d] = v;
    s2 = fmaf(v, v, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  __syncwarp();
  return s2;
}

// ------------------------------------------------------------------------------------
// exact re-rank of rows whose approximate order is ambiguous
// ------------------------------------------------------------------------------------
// One warp per listed row; lane c (and c + 32) owns candidate c of the row's <= TL candidates: exact fp32
// distance with the arithmetic of knn_exact.cu, rank by counting over (distance, id), dilated pick.  The
// row goes on to the brute-force fix-up when a key outside the candidate list could still belong to the
// k*d nearest.  List entry: [row, np, approx bound of the keys outside, ids[TL], approx values[TL]].
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const int* __restrict__ rr_count, const int* __restrict__ rr_list, int rr_cap, int TL,
                  RawFeat xf, const float* __restrict__ yhat, const float* __restrict__ ysq,
                  const float* __restrict__ relpos, int32_t* __restrict__ idx_out,
                  int* fix_count, int* fix_rows, unsigned int* stats, int N, int M, int D, int k, int dilation,
                  float delta) {
  extern __shared__ __align__(16) float xrow_s[];      // [8 warps][D4 * 4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xr = xrow_s + warp * ((D + 3) / 4 * 4);
  const int total = min(*rr_count, rr_cap);
  const int kd = k * dilation;
  const int stride = 2 * TL + 3;
  for (int e = blockIdx.x * (blockDim.x >> 5) + warp; e < total; e += gridDim.x * (blockDim.x >> 5)) {
    const int* ent = rr_list + (size_t)e * stride;
    const int row = ent[0], np = ent[1];
    const float a_last = __int_as_float(ent[2]);
    const int p = row / N, n = row - p * N;
    __syncwarp();
    const float xs = normalise_row_warp(xf, p, n, D, xr, lane);
    const float* relrow = relpos ? relpos + (size_t)n * M : nullptr;
    float ev[2];
    int id[2];
    float maxerr = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = lane + 32 * h;
      ev[h] = INFINITY;
      id[h] = 0x7fffffff;
      if (c < np) {
        id[h] = ent[3 + c];
        ev[h] = exact_dist(xr, yhat + ((size_t)p * M + id[h]) * D, D, xs, ysq[(size_t)p * M + id[h]], relrow, id[h]);
        maxerr = fmaxf(maxerr, fabsf((ev[h] - xs) - __int_as_float(ent[3 + TL + c])));
      }
    }
    int rank[2] = {0, 0};
    for (int j = 0; j < np; ++j) {
      const float vj = __shfl_sync(0xffffffffu, j < 32 ? ev[0] : ev[1], j & 31);
      const int ij = __shfl_sync(0xffffffffu, j < 32 ? id[0] : id[1], j & 31);
#pragma unroll
      for (int h = 0; h < 2; ++h) rank[h] += (vj < ev[h]) || (vj == ev[h] && ij < id[h]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxerr = fmaxf(maxerr, __shfl_xor_sync(0xffffffffu, maxerr, o));
    if (lane == 0) atomicMax(stats + 1, __float_as_uint(maxerr));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (lane + 32 * h >= np) continue;
      if (rank[h] < kd && rank[h] % dilation == 0) idx_out[(size_t)row * k + rank[h] / dilation] = id[h];
      // the kd-th nearest candidate decides whether a key outside the list could still matter
      // (every key is in the list when np == M: nothing outside to worry about)
      if (rank[h] == kd - 1 && np < M && a_last - delta <= (ev[h] - xs) + delta)
        fix_rows[atomicAdd(fix_count, 1)] = row;
    }
  }
}

// ------------------------------------------------------------------------------------
// exact fix-up for rows whose candidate set could not be certified (rare)
// ------------------------------------------------------------------------------------
// Per listed row, one CTA: exact distances to all M keys, then the k*d nearest by (distance, id).  Selection:
// a 1024-bin histogram over [min, max] (the bin index is monotone in the distance) locates the bin that holds
// the k*d-th nearest; the keys up to that bin -- a superset of the answer, every other key is strictly farther --
// are compacted and ranked by counting among themselves.  When they exceed the candidate slots (a pile of tied
// distances) the counting runs over all M keys (still exact).
constexpr int kFixBins = 1024;
constexpr int kFixCand = 2048;                 // candidate slots (keys up to the selected bin)
constexpr int kFixTile = 8192;                 // keys whose distances live in shared memory at a time
constexpr int kFixThreads = 1024;              // a row's keys over 32 warps: the exact distance is a serial D-long chain of L2 loads per key
__global__ void __launch_bounds__(kFixThreads)
knn_fixup_kernel(const int* __restrict__ count, const int* __restrict__ rows, RawFeat xf,
                 const float* __restrict__ yhat, const float* __restrict__ ysq,
                 const float* __restrict__ relpos, int32_t* __restrict__ idx_out, float* __restrict__ dist_g,
                 int N, int M, int D, int k, int dilation) {
  extern __shared__ __align__(16) float fix_s[];   // [D4*4] query row, [Ms] distances, [cap] candidate distances, [cap] candidate ids
  const int Ms = M <= kFixTile ? M : 0;            // more keys than that: the distances go to a global scratch row
  const int cap = M < kFixCand ? M : kFixCand;
  float* xr = fix_s;
  float* dist_s = fix_s + (D + 3) / 4 * 4;
  float* cand_v = dist_s + Ms;
  int* cand_i = reinterpret_cast<int*>(cand_v + cap);
  __shared__ int hist[kFixBins];
  __shared__ float red_lo[kFixThreads / 32], red_hi[kFixThreads / 32];
  __shared__ float xs_s;
  __shared__ int bin_sel, ncand, nsel_s;
  const int total = *count;
  const int kd = k * dilation;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* dist = Ms > 0 ? dist_s : dist_g + (size_t)blockIdx.x * M;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int id = rows[item];
    const int p = id / N, n = id - p * N;
    if (warp == 0) {
      const float s = normalise_row_warp(xf, p, n, D, xr, lane);
      if (lane == 0) xs_s = s;
    }
    for (int q = threadIdx.x; q < kFixBins; q += blockDim.x) hist[q] = 0;
    if (threadIdx.x == 0) ncand = 0;
    __syncthreads();
    const float xs = xs_s;
    const float* relrow = relpos ? relpos + (size_t)n * M : nullptr;
    float lo = INFINITY, hi = -INFINITY;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      const float v = exact_dist(xr, yhat + ((size_t)p * M + m) * D, D, xs, ysq[(size_t)p * M + m], relrow, m);
      dist[m] = v;
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kFixThreads / 32; ++w) { lo = fminf(lo, red_lo[w]); hi = fmaxf(hi, red_hi[w]); }
    const float scale = hi > lo ? (float)kFixBins / (hi - lo) : 0.f;
    auto bin_of = [&](float v) {
      const int q = (int)((v - lo) * scale);
      return q < 0 ? 0 : (q < kFixBins ? q : kFixBins - 1);
    };
    for (int m = threadIdx.x; m < M; m += blockDim.x) atomicAdd(&hist[bin_of(dist[m])], 1);
    __syncthreads();
    if (warp == 0) {                           // first bin whose prefix count reaches k*d
      constexpr int PER = kFixBins / 32;
      int part = 0;
      for (int q = 0; q < PER; ++q) part += hist[lane * PER + q];
      int incl = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned reach = __ballot_sync(0xffffffffu, incl >= kd);
      const int first = reach ? __ffs(reach) - 1 : 31;
      if (lane == first) {
        int run = incl - part, q = 0;
        for (; q < PER; ++q) {                  // keys up to and including bin q of this lane
          run += hist[lane * PER + q];
          if (run >= kd) break;
        }
        if (q == PER) q = PER - 1;
        bin_sel = reach ? lane * PER + q : kFixBins - 1;
        nsel_s = reach ? run : M;
      }
    }
    __syncthreads();
    const int nsel = nsel_s;
    const int bsel = bin_sel;
    if (nsel <= cap) {
      for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float v = dist[m];
        if (bin_of(v) <= bsel) {
          const int pos = atomicAdd(&ncand, 1);
          if (pos < cap) { cand_v[pos] = v; cand_i[pos] = m; }
        }
      }
      __syncthreads();
      const int nc = ncand < cap ? ncand : cap;
      for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        const float v = cand_v[c];
        const int m = cand_i[c];
        int rank = 0;
        for (int j = 0; j < nc; ++j) {
          const float u = cand_v[j];
          rank += (u < v) || (u == v && cand_i[j] < m);
        }
        if (rank < kd && rank % dilation == 0) idx_out[(size_t)id * k + rank / dilation] = m;
      }
    } else {                                   // (nearly) all distances in one bin: count over every key
      for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float v = dist[m];
        if (bin_of(v) > bsel) continue;
        int rank = 0;
        for (int j = 0; j < M && rank < kd; ++j) {
          const float u = dist[j];
          rank += (u < v) || (u == v && j < m);
        }
        if (rank < kd && rank % dilation == 0) idx_out[(size_t)id * k + rank / dilation] = m;
      }
    }
    __syncthreads();
  }
}
constexpr int kFixBlocks = 148;

struct TcWorkspace {
  __half* a_op; __half* b_op; int* fix_count; int* fix_rows; unsigned int* stats;
  int* rr_count; int* rr_list; int rr_cap; float* fix_dist;
  float2* cand; int* cand_count; float* cand_thr; int cand_slots; size_t bytes;
};

// rows the re-rank list can hold (ambiguous rows are a few per thousand; beyond the cap they take the
// brute-force fix-up, which is always correct)
static int rerank_cap(int P, int N) {
  const long long rows = (long long)P * N;
  const long long cap = rows / 4 > 4096 ? rows / 4 : 4096;
  return (int)(cap < rows ? cap : rows);
}

TcWorkspace carve_tc(void* base, const Plan& pl, int P, int N, int M, int T) {
  TcWorkspace w;
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t n) { void* p = b ? b + off : nullptr; off += align_up(n, 256); return p; };
  w.a_op = static_cast<__half*>(take(pl.a_op_bytes));
  w.b_op = static_cast<__half*>(take(pl.b_op_bytes));
  w.fix_count = static_cast<int*>(take(256));
  w.stats = reinterpret_cast<unsigned int*>(w.fix_count ? w.fix_count + 4 : nullptr);
  w.rr_count = w.fix_count ? w.fix_count + 8 : nullptr;
  w.fix_rows = static_cast<int*>(take(sizeof(int) * (size_t)P * N));
  w.rr_cap = rerank_cap(P, N);
  w.rr_list = static_cast<int*>(take(sizeof(int) * (size_t)w.rr_cap * (2 * (T + 3) + 3)));
  w.fix_dist = static_cast<float*>(take(M > kFixTile ? sizeof(float) * (size_t)kFixBlocks * M : 0));
  {   // candidate hand-over of the select kernel: per item (pl.QI items of rows_per_item rows per problem)
    const int rows_per_item = (pl.geom == 1 ? GeomB::ROWS : GeomA::ROWS);
    const size_t item_rows = (size_t)P * pl.QI * rows_per_item;
    w.cand_slots = T + (T / 2 + 1 > 9 ? T / 2 + 1 : 9);
    w.cand = static_cast<float2*>(take(sizeof(float2) * item_rows * w.cand_slots));
    w.cand_count = static_cast<int*>(take(sizeof(int) * item_rows));
    w.cand_thr = static_cast<float*>(take(sizeof(float) * item_rows));
  }
  w.bytes = off;
  return w;
}

}  // namespace

static int t_bucket(int T) { return T <= 11 ? 11 : T <= 20 ? 20 : T <= 29 ? 29 : 38; }

bool knn_tc_supported(int N, int M, int D, int k, int dilation, int dtype) {
  (void)dtype;
  const int kd = k * dilation;
  if (kd + 2 > MAX_T || kd > M) return false;
  if (N < 1 || M < 1 || M > 65000) return false;   // key ids travel in 16 bits of a log entry
  Plan pl = make_plan(1, N, M, D, t_bucket(kd + 2));
  return pl.ok;
}

// AUTO picks the tensor-core path only when the threshold sweep has enough key groups to work with
// (tiny key sets would send every row to the brute-force fix-up: correct, but the exact kernel is faster).
bool knn_tc_preferred(int N, int M, int D, int k, int dilation, int dtype) {
  return knn_tc_supported(N, M, D, k, dilation, dtype) && M / 3 >= 2 * t_bucket(k * dilation + 2);
}

size_t knn_tc_workspace_bytes(int P, int N, int M, int D, int k, int dilation, int dtype) {
  (void)dtype;
  Plan pl = make_plan(P, N, M, D, t_bucket(k * dilation + 2));
  if (!pl.ok) return 0;
  return carve_tc(nullptr, pl, P, N, M, t_bucket(k * dilation + 2)).bytes;
}

// row-per-thread path: D in {20, 40, 80}, rows 16-byte aligned
template <typename T, int D>
static int launch_prepare_rows(const KnnWorkspace& w, const TcWorkspace& t, const Plan& pl, const T* x, int64_t x_sb,
                               int64_t x_sn, const T* y, int64_t y_sb, int64_t y_sn, int P, int G, int N, int M,
                               bool self_keys, cudaStream_t stream) {
  const int bn = pl.geom == 1 ? GeomB::BN : GeomA::BN, bnp = pl.geom == 1 ? GeomB::BNP : GeomA::BNP;
  {
    static_assert(BM == PREP_THREADS, "a block of the query kernel assembles exactly one operand tile");
    dim3 grid((pl.QTP * BM + PREP_THREADS - 1) / PREP_THREADS, P);
    const size_t stage_bytes = (size_t)BM * k_padded(D) * 2;
    static std::atomic<uint64_t> configured{0};
    configure_once_per_device(configured, [] {
      cudaFuncSetAttribute(tc_prepare_rows_kernel<T, D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    tc_prepare_rows_kernel<T, D, false><<<grid, PREP_THREADS, stage_bytes, stream>>>(x, x_sb, x_sn, nullptr, nullptr, t.a_op,
                                                                                     G, N, pl.KC, pl.QTP, BM, BM, 0);
    GKG_CHECK_LAUNCH("tc_prepare_rows_kernel<query>");
  }
  {
    dim3 grid((pl.KT * bnp + PREP_THREADS - 1) / PREP_THREADS, P);
    const T* src = self_keys ? x : y;
    tc_prepare_rows_kernel<T, D, true><<<grid, PREP_THREADS, 0, stream>>>(
        src, self_keys ? x_sb : y_sb, self_keys ? x_sn : y_sn, w.yhat, w.ysq, t.b_op, G, M, pl.KC, pl.KT, bnp, bn, 1);
    GKG_CHECK_LAUNCH("tc_prepare_rows_kernel<key>");
  }
  return GKG_OK;
}

template <typename T>
static bool rows_path_ok(const void* x, int64_t x_sb, int64_t x_sn, const void* y, int64_t y_sb, int64_t y_sn,
                         int D, bool self_keys) {
  const int ae = (int)(16 / sizeof(T));
  auto ok = [&](const void* p, int64_t sb, int64_t sn) {
    return ((uintptr_t)p % 16) == 0 && sb % ae == 0 && sn % ae == 0;
  };
  return D % ae == 0 && ok(x, x_sb, x_sn) && (self_keys || ok(y, y_sb, y_sn));
}

template <typename T>
static int launch_prepare_typed(const KnnWorkspace& w, const TcWorkspace& t, const Plan& pl, const void* x,
                                int64_t x_sb, int64_t x_sn, const void* y, int64_t y_sb, int64_t y_sn, int P,
                                int G, int N, int M, int D, bool self_keys, cudaStream_t stream) {
  if (rows_path_ok<T>(x, x_sb, x_sn, y, y_sb, y_sn, D, self_keys)) {
    const T* xt = static_cast<const T*>(x);
    const T* yt = static_cast<const T*>(y);
    switch (D) {
      case 20: return launch_prepare_rows<T, 20>(w, t, pl, xt, x_sb, x_sn, yt, y_sb, y_sn, P, G, N, M, self_keys, stream);
      case 40: return launch_prepare_rows<T, 40>(w, t, pl, xt, x_sb, x_sn, yt, y_sb, y_sn, P, G, N, M, self_keys, stream);
      case 80: return launch_prepare_rows<T, 80>(w, t, pl, xt, x_sb, x_sn, yt, y_sb, y_sn, P, G, N, M, self_keys, stream);
      default: break;
    }
  }
  const size_t smem = sizeof(float) * ((size_t)PREP_ROWS * prep_stride(D) + PREP_ROWS);
  const int bn = pl.geom == 1 ? GeomB::BN : GeomA::BN, bnp = pl.geom == 1 ? GeomB::BNP : GeomA::BNP;
  {
    dim3 grid((pl.QTP * BM + PREP_ROWS - 1) / PREP_ROWS, P);
    static std::atomic<uint64_t> configured{0};
    configure_once_per_device(configured, [] {   // wide groups (D > 380) need more than 48 KB of dynamic shared memory
      cudaFuncSetAttribute(tc_prepare_kernel<T, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(tc_prepare_kernel<T, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    auto kq = D == 200 ? tc_prepare_kernel<T, false, 200> : D == 320 ? tc_prepare_kernel<T, false, 320> : tc_prepare_kernel<T, false, 0>;
    kq<<<grid, 256, smem, stream>>>(static_cast<const T*>(x), x_sb, x_sn, nullptr, nullptr, t.a_op, G, N, D, pl.KP, pl.KC,
                                    pl.QTP, BM, BM, 0, pl.split ? pl.PA : 0);
    GKG_CHECK_LAUNCH("tc_prepare_kernel<query>");
  }
  {
    dim3 grid((pl.KT * bnp + PREP_ROWS - 1) / PREP_ROWS, P);
    const T* src = static_cast<const T*>(self_keys ? x : y);
    auto kk = D == 200 ? tc_prepare_kernel<T, true, 200> : D == 320 ? tc_prepare_kernel<T, true, 320> : tc_prepare_kernel<T, true, 0>;
    kk<<<grid, 256, smem, stream>>>(src, self_keys ? x_sb : y_sb, self_keys ? x_sn : y_sn, w.yhat, w.ysq, t.b_op, G, M, D,
                                    pl.KP, pl.KC, pl.KT, bnp, bn, 1, pl.split ? pl.PA : 0);
    GKG_CHECK_LAUNCH("tc_prepare_kernel<key>");
  }
  return GKG_OK;
}

int launch_knn_tc_prepare(const KnnWorkspace& w, void* extra_ws, const void* x, int64_t x_sb, int64_t x_sn,
                          const void* y, int64_t y_sb, int64_t y_sn, int dtype, int P, int G, int N, int M,
                          int D, int k, int dilation, bool self_keys, cudaStream_t stream) {
  Plan pl = make_plan(P, N, M, D, t_bucket(k * dilation + 2));
  GKG_CHECK_ARG(pl.ok, "knn_tc: no tiling for D=%d", D);
  GKG_CHECK_ARG(P <= 65535, "knn_tc: B*G=%d > 65535", P);
  TcWorkspace t = carve_tc(extra_ws, pl, P, N, M, t_bucket(k * dilation + 2));
  if (dtype == GKG_F32)
    return launch_prepare_typed<float>(w, t, pl, x, x_sb, x_sn, y, y_sb, y_sn, P, G, N, M, D, self_keys, stream);
  return launch_prepare_typed<__nv_bfloat16>(w, t, pl, x, x_sb, x_sn, y, y_sb, y_sn, P, G, N, M, D, self_keys,
                                             stream);
}

int launch_knn_tc(const KnnWorkspace& w, void* extra_ws, const void* x, int64_t x_sb, int64_t x_sn, int dtype, int G,
                  const float* relpos, const SepBias& sep, int32_t* idx_out, int P, int N, int M, int D, int k,
                  int dilation, const KnnDebug* dbg, cudaStream_t stream) {
  const int flags = dbg ? dbg->flags : 0;
  Plan pl = make_plan(P, N, M, D, t_bucket(k * dilation + 2));
  GKG_CHECK_ARG(pl.ok, "knn_tc: no tiling for D=%d", D);
  const int T = t_bucket(k * dilation + 2);
  TcWorkspace t = carve_tc(extra_ws, pl, P, N, M, T);
  cudaError_t e = cudaMemsetAsync(t.fix_count, 0, 256, stream);
  if (e != cudaSuccess) {
    set_error("knn_tc: memset: %s", cudaGetErrorString(e));
    return GKG_ECUDA;
  }
  count_launch();
  TcParams prm{};
  prm.a_op = t.a_op; prm.b_op = t.b_op;
  prm.yhat = w.yhat; prm.ysq = w.ysq;
  prm.relpos = relpos; prm.idx_out = idx_out;
  prm.fix_count = t.fix_count; prm.fix_rows = t.fix_rows; prm.stats = t.stats;
  prm.rr_count = t.rr_count; prm.rr_list = t.rr_list; prm.rr_cap = t.rr_cap;
  prm.cand = t.cand; prm.cand_count = t.cand_count; prm.cand_thr = t.cand_thr; prm.cand_slots = t.cand_slots;
  prm.dbg_dist = dbg ? dbg->dist : nullptr;
  prm.P = P; prm.N = N; prm.M = M; prm.D = D; prm.k = k; prm.dilation = dilation; prm.kd = k * dilation;
  prm.KP = pl.KP; prm.PA = pl.PA; prm.KC = pl.KC; prm.NKB = pl.NKB; prm.NKBA = pl.NKBA; prm.NA = pl.NA; prm.NS = pl.NS; prm.QT = pl.QT;
  prm.QI = pl.QI; prm.QTP = pl.QTP; prm.KT = pl.KT;
  prm.split = pl.split; prm.KS1 = pl.KS1; prm.KSL = pl.KSL;
  prm.a_tile_bytes = pl.a_tile_bytes; prm.a_res_bytes = pl.a_res_bytes; prm.b_block_bytes = pl.b_block_bytes; prm.stage_bytes = pl.stage_bytes;
  prm.force_rerank = flags;
  prm.delta = pl.split ? tc_delta_steps(pl.KS1 + 2 * pl.KSL) : tc_delta(pl.KP);
  // the separable form of the bias is a hint the caller verified to within kSepBiasTol of the dense table the exact
  // re-rank reads (include/gkg_abi.h): that residual is part of the approximation error the gap test must cover
  constexpr float kSepBiasTol = 5e-7f;
  if (relpos != nullptr && sep.a != nullptr && sep.b != nullptr) prm.delta += kSepBiasTol;
  int ga = (M / 18 >= 3 * (T - 1)) ? 18 : (M / 6 >= 3 * (T - 1)) ? 6 : 3;
  if (dbg && (dbg->ga == 18 || dbg->ga == 6 || dbg->ga == 3)) ga = dbg->ga;
  prm.sep_a = sep.a; prm.sep_b = sep.b; prm.grid_w = sep.grid_w > 0 ? sep.grid_w : 1;
  prm.sep_mh = sep.kw > 0 ? M / sep.kw : 1;
  const int bn = pl.geom == 1 ? GeomB::BN : GeomA::BN;
  const int sepw = pl.geom == 1 ? GeomB::SEPW : GeomA::SEPW;
  prm.sep_mhp = sep.kw > 0 ? pl.KT * bn / sep.kw : 1;
  int bias = relpos != nullptr ? 1 : 0;
  if (bias && sep.a != nullptr && sep.b != nullptr && (sep.kw == 9 || sep.kw == 18 || sep.kw == 36) &&
      sep.grid_w > 0 && N % sep.grid_w == 0 && M % sep.kw == 0 &&
      (32 / sep.grid_w + 2) * (pl.KT * bn / sep.kw) <= sepw)
    bias = sep.kw;
  int rc;
  switch (bias) {
    case 0: rc = tc::launch_select<0>(prm, pl, T, ga, stream); break;
    case 9: rc = tc::launch_select<9>(prm, pl, T, ga, stream); break;
    case 18: rc = tc::launch_select<18>(prm, pl, T, ga, stream); break;
    case 36: rc = tc::launch_select<36>(prm, pl, T, ga, stream); break;
    default: rc = tc::launch_select<1>(prm, pl, T, ga, stream); break;
  }
  if (rc != GKG_OK) return rc;
  {
    const int rows_per_item = (pl.geom == 1 ? GeomB::ROWS : GeomA::ROWS);
    const int row_sets = rows_per_item / BM;
    const long long total_rows = (long long)P * pl.QI * rows_per_item;
    const unsigned blocks = (unsigned)((total_rows + 255) / 256);
#define GKG_FINALIZE(LL)                                                                                          \
    knn_finalize_kernel<LL><<<blocks, 256, 0, stream>>>(t.cand, t.cand_count, t.cand_thr, t.cand_slots, rows_per_item, \
        row_sets, pl.QI, idx_out, t.rr_count, t.rr_list, t.rr_cap, t.fix_count, t.fix_rows, t.stats, total_rows, N, M,  \
        k, dilation, prm.delta, flags)
    if (T == 11) GKG_FINALIZE(14); else if (T == 20) GKG_FINALIZE(23); else if (T == 29) GKG_FINALIZE(32); else GKG_FINALIZE(41);
#undef GKG_FINALIZE
    GKG_CHECK_LAUNCH("knn_finalize_kernel");
  }
  RawFeat xf{x, x_sb, x_sn, dtype, G};
  {
    const int blocks = (t.rr_cap + 7) / 8 < 148 * 4 ? (t.rr_cap + 7) / 8 : 148 * 4;
    const size_t rsmem = sizeof(float) * 8 * (size_t)((D + 3) / 4 * 4);
    knn_rerank_kernel<<<blocks, 256, rsmem, stream>>>(t.rr_count, t.rr_list, t.rr_cap, T + 3, xf, w.yhat, w.ysq,
                                                      relpos, idx_out, t.fix_count, t.fix_rows, t.stats, N, M, D, k,
                                                      dilation, prm.delta);
    GKG_CHECK_LAUNCH("knn_rerank_kernel");
  }
  {
    const size_t fsmem = sizeof(float) * ((size_t)((D + 3) / 4 * 4) + (M <= kFixTile ? (size_t)M : 0) +
                                          2 * (size_t)(M < kFixCand ? M : kFixCand));
    static std::atomic<uint64_t> configured{0};
    configure_once_per_device(configured, [] {
      cudaFuncSetAttribute(knn_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    });
    knn_fixup_kernel<<<kFixBlocks, kFixThreads, fsmem, stream>>>(t.fix_count, t.fix_rows, xf, w.yhat, w.ysq, relpos, idx_out,
                                                         t.fix_dist, N, M, D, k, dilation);
    GKG_CHECK_LAUNCH("knn_fixup_kernel");
  }
  if (dbg && dbg->stats_out) {   // tests only: expose the counters [fix-up rows, ambiguous rows, max error bits]
    cudaStreamSynchronize(stream);
    unsigned int h[12];
    cudaMemcpy(h, t.fix_count, sizeof(h), cudaMemcpyDeviceToHost);
    dbg->stats_out[0] = h[0];
    dbg->stats_out[1] = h[4];
    dbg->stats_out[2] = h[5];
  }
  return GKG_OK;
}

}  // namespace gkg


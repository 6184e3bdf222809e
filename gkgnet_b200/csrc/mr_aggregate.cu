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
#include "knn_tc.cuh"
#include <stdlib.h>

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
        // strictly greater: the first maximum wins ties; a NaN neighbour is taken once and then kept (torch.max propagates NaN)
        if (f > best[e] || (f != f && best[e] == best[e])) { best[e] = f; arg[e] = j; }
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

// DET: the scattered sum runs in 64-bit fixed point (value * scale, scale a power of two read from device memory):
// integer addition is associative, so the result does not depend on the order in which the atomics land.
template <typename T, int VEC, bool DET>
__global__ void __launch_bounds__(256)
mr_aggregate_bwd_kernel(const T* __restrict__ gout, const int32_t* __restrict__ idx,
                        const uint8_t* __restrict__ argmax, T* __restrict__ gx,
                        float* __restrict__ gy, const float* __restrict__ scale_ptr, int G, int N, int M, int D, int k,
                        long long total_chunks) {
  const float scale = DET ? __ldg(scale_ptr) : 1.f;
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
    float* gyb = gy + (DET ? 2 : 1) * ((b * M) * (long long)C + c0);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      o.v[e] = from_f32<T>(g_self[e] - g_rel[e]);
      const int m = __ldg(ip + am.v[e]);
      if (DET)
        atomicAdd(reinterpret_cast<unsigned long long*>(gyb) + (long long)m * C + e,
                  (unsigned long long)__float2ll_rn(g_rel[e] * scale));
      else
        atomicAdd(gyb + (long long)m * C + e, g_rel[e]);
    }
    *reinterpret_cast<P*>(gx + bn * C + c0) = o;
  }
}

// ---- bf16 fast path: 8 channels (16 bytes) per thread, packed bf16x2 arithmetic -------------------
// The maximum over the K neighbours is taken with HMNMX2 (exact: max selects one of its inputs); the
// arg-max is recovered afterwards from equality masks (smallest j wins, like the scalar kernel), so the
// per-element convert / compare / select chain of the generic kernel disappears.  m = max - x is formed
// in fp32 and rounded once, exactly as the generic kernel (and the reference under autocast) does.
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
  // HMNMX2.NAN: a NaN neighbour propagates into the maximum, as torch.max does
  __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf2_eq_mask(uint32_t a, uint32_t b) {
  return __heq2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}
// (x_lo, x_hi), (best_lo, best_hi) -> interleaved (x_lo, m_lo), (x_hi, m_hi)
__device__ __forceinline__ void bf2_emit(uint32_t x, uint32_t best, uint32_t& o0, uint32_t& o1) {
  const float xl = __uint_as_float(x << 16), xh = __uint_as_float(x & 0xffff0000u);
  const float bl = __uint_as_float(best << 16), bh = __uint_as_float(best & 0xffff0000u);
  __nv_bfloat162 m = __floats2bfloat162_rn(bl - xl, bh - xh);
  const uint32_t mu = *reinterpret_cast<uint32_t*>(&m);
  o0 = __byte_perm(x, mu, 0x5410);
  o1 = __byte_perm(x, mu, 0x7632);
}

template <int K, bool ARG>
__device__ __forceinline__ void mr_chunk_bf16(const uint4 xv, const uint4 (&yv)[K], __nv_bfloat16* __restrict__ op,
                                              uint8_t* __restrict__ ap) {
  uint32_t b0 = yv[0].x, b1 = yv[0].y, b2 = yv[0].z, b3 = yv[0].w;
#pragma unroll
  for (int j = 1; j < K; ++j) {
    b0 = bf2_max(b0, yv[j].x); b1 = bf2_max(b1, yv[j].y); b2 = bf2_max(b2, yv[j].z); b3 = bf2_max(b3, yv[j].w);
  }
  uint4 o0, o1;
  bf2_emit(xv.x, b0, o0.x, o0.y); bf2_emit(xv.y, b1, o0.z, o0.w);
  bf2_emit(xv.z, b2, o1.x, o1.y); bf2_emit(xv.w, b3, o1.z, o1.w);
  reinterpret_cast<uint4*>(op)[0] = o0;
  reinterpret_cast<uint4*>(op)[1] = o1;
  if (ARG) {
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;       // two 16-bit neighbour slots per register
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {
      const uint32_t jj = (uint32_t)j * 0x00010001u;
      uint32_t m;
      m = bf2_eq_mask(yv[j].x, b0); a0 = (a0 & ~m) | (jj & m);
      m = bf2_eq_mask(yv[j].y, b1); a1 = (a1 & ~m) | (jj & m);
      m = bf2_eq_mask(yv[j].z, b2); a2 = (a2 & ~m) | (jj & m);
      m = bf2_eq_mask(yv[j].w, b3); a3 = (a3 & ~m) | (jj & m);
    }
    uint2 a;
    a.x = __byte_perm(a0, a1, 0x6420);
    a.y = __byte_perm(a2, a3, 0x6420);
    *reinterpret_cast<uint2*>(ap) = a;
  }
}

template <int K, bool ARG>
__global__ void __launch_bounds__(256)
mr_aggregate_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_sb, int64_t x_sn,
                             const __nv_bfloat16* __restrict__ y, int64_t y_sb, int64_t y_sn,
                             const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                             uint8_t* __restrict__ argmax, int G, int N, int D, long long total_chunks) {
  const int C = G * D;
  const int chunks_per_node = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % chunks_per_node);
    const long long bn = i / chunks_per_node;
    const int n = (int)(bn % N);
    const long long b = bn / N;
    const int c0 = cc << 3;
    const int g = c0 / D;
    const int32_t* ip = idx + ((b * G + g) * N + n) * (long long)K;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + b * x_sb + (long long)n * x_sn + c0);
    const __nv_bfloat16* ybase = y + b * y_sb + c0;
    uint4 yv[K];
#pragma unroll
    for (int j = 0; j < K; ++j) yv[j] = *reinterpret_cast<const uint4*>(ybase + (long long)__ldg(ip + j) * y_sn);
    mr_chunk_bf16<K, ARG>(xv, yv, out + (bn * C + c0) * 2, ARG ? argmax + bn * C + c0 : nullptr);
  }
}

// ---- bf16 fast path with the key rows staged in shared memory ---------------------------------------
// The gather reads every key row ~N*k/M times (144 at stage 1): served from L1/L2 it is the l1tex
// pipe, not HBM, that bounds the kernel.  Here a CTA owns a contiguous range of (image, channel
// slice, node) work, copies the slice's keys (M x CS bf16 <= 110 KB, two CTAs per SM) into shared
// memory once per (image, slice) it touches and gathers from there; global memory then only sees
// the streaming traffic (x, idx in; out, argmax out).
constexpr int kAggSmemThreads = 512;
constexpr size_t kAggSmemKeysMax = 102 * 1024;   // + 2 idx tiles (<= 10 KB) per CTA, two CTAs per SM

template <int K, bool ARG>
__global__ void __launch_bounds__(kAggSmemThreads, 2)
mr_aggregate_fwd_bf16_smem_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_sb, int64_t x_sn,
                                  const __nv_bfloat16* __restrict__ y, int64_t y_sb, int64_t y_sn,
                                  const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                  uint8_t* __restrict__ argmax, int B, int G, int N, int M, int D, int CS) {
  extern __shared__ __align__(16) uint8_t agg_smem[];          // [M][CS] bf16 keys, then 2 idx tiles
  const int C = G * D;
  const int NSL = C / CS;                                       // channel slices per image
  const int CPN = CS >> 3;                                      // 16-byte chunks per node slice
  const int npp = kAggSmemThreads / CPN;                        // nodes per pass of the CTA
  const int node_l = threadIdx.x / CPN, ch = threadIdx.x - node_l * CPN;
  const bool active = node_l < npp;
  // CTA c works on slice c % NSL of node-row range c / NSL: the CTAs that need the same x rows run
  // side by side, so x comes from HBM once and from L2 afterwards
  const int slice = blockIdx.x % NSL;
  const int range = blockIdx.x / NSL, nranges = gridDim.x / NSL;
  const long long R = (long long)B * N;                         // node rows in total
  const long long per = (R + nranges - 1) / nranges;
  long long r = (long long)range * per;
  const long long r_end = r + per < R ? r + per : R;
  const int c0 = slice * CS, g = c0 / D, cc = c0 + ch * 8;
  const uint32_t keys_s = (uint32_t)__cvta_generic_to_shared(agg_smem);
  const uint32_t row_bytes = (uint32_t)CS * 2;
  const uint32_t tile_s = keys_s + (uint32_t)M * row_bytes;     // 2 x [npp][K] int32
  const uint32_t tile_bytes = (uint32_t)npp * K * 4;
  if (range >= nranges) return;
  while (r < r_end) {
    const long long b = r / N;
    const int n_lo = (int)(r - b * N);
    const int n_hi = (int)((long long)n_lo + (r_end - r) < (long long)N ? n_lo + (r_end - r) : N);
    const int32_t* ibase = idx + ((b * G + g) * N) * (long long)K;
    auto fetch_tile = [&](int node0, int buf) {                 // idx rows of nodes [node0, node0 + npp) -> smem
      const int cnt = (min(n_hi, node0 + npp) - node0) * K;
      const int32_t* src = ibase + (long long)node0 * K;
      for (int q = threadIdx.x; q < cnt; q += kAggSmemThreads)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile_s + buf * tile_bytes + q * 4), "l"(src + q)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_x = [&](int node) {
      return node < n_hi && active ? *reinterpret_cast<const uint4*>(x + b * x_sb + (long long)node * x_sn + cc)
                                   : make_uint4(0, 0, 0, 0);
    };
    __syncthreads();                                            // gathers of the previous image are done
    {
      const __nv_bfloat16* src = y + b * y_sb + c0;
      const int pieces = M * CPN;
      for (int q = threadIdx.x; q < pieces; q += kAggSmemThreads) {
        const int m = q / CPN, qc = q - m * CPN;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(keys_s + (uint32_t)m * row_bytes + qc * 16),
                     "l"(src + (long long)m * y_sn + qc * 8) : "memory");
      }
    }
    fetch_tile(n_lo, 0);                                        // same group as the keys
    uint4 xv = load_x(n_lo + node_l);
    int buf = 0;
    for (int node0 = n_lo; node0 < n_hi; node0 += npp, buf ^= 1) {
      __syncthreads();                                          // everyone is done with tile buf ^ 1
      if (node0 + npp < n_hi) fetch_tile(node0 + npp, buf ^ 1); else asm volatile("cp.async.commit_group;" ::: "memory");
      const uint4 xn = load_x(node0 + npp + node_l);
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // keys + tile `buf` have landed
      __syncthreads();
      const int node = node0 + node_l;
      if (active && node < n_hi) {
        const long long bn = b * N + node;
        const uint32_t ip = tile_s + buf * tile_bytes + (uint32_t)node_l * K * 4;
        uint4 yv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
          uint32_t m;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(ip + j * 4));
          const uint32_t a = keys_s + m * row_bytes + ch * 16;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(yv[j].x), "=r"(yv[j].y), "=r"(yv[j].z), "=r"(yv[j].w) : "r"(a));
        }
        mr_chunk_bf16<K, ARG>(xv, yv, out + (bn * C + cc) * 2, ARG ? argmax + bn * C + cc : nullptr);
      }
      xv = xn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    r += n_hi - n_lo;
  }
}

// ---- bf16 fast path, warp-autonomous form --------------------------------------------------------
// Same idea (keys of one image slice resident in shared memory), without the CTA-wide idx tiles and
// their two barriers per tile: a warp owns NPW = 32 / CPN consecutive nodes per pass (CPN = 16-byte
// chunks per node slice), the SUB lanes that share a neighbour list (same node, same channel group)
// load it with coalesced LDG.32 one pass ahead (two ids packed per register) and hand the ids round
// with SHFL, the running maximum and its arg-max are kept packed (HMNMX2 / HSET2 mask / LOP3), the
// 32 output bytes of a lane leave in one 256-bit store.  With CS = 80 at stage 1 a CTA covers whole
// x / out rows (dense 128-byte lines: the 40-channel slices of the kernel above touched twice the
// lines per request), one 1024-thread CTA per SM, the only barriers are at image boundaries.
//
// Bank conflicts: a 128-bit LDS is served a quarter warp (8 lanes, 128 bytes) at a time.  With
// 160-byte key rows the 8 lanes straddle two random rows and collide 7 times out of 8.  For CS = 80
// the rows are therefore split: chunks 0-7 live in a main array of pitch 128 bytes (chunk c always
// in 16-byte bank group c), chunks 8-9 in an aux array of pitch 32 bytes.  Lanes 8i..8i+7 take
// chunks 0-7 of node i (always conflict free, whatever the rows), lanes 24+2i, 25+2i its chunks 8-9
// (three random pairs over four positions: 1.7 wavefronts on average): 4.7 wavefronts per gather
// instruction instead of 7.4, same shared-memory footprint.
constexpr size_t kAggWarpSmemMax = 226 * 1024;
#ifndef GKG_AGG_PF
#define GKG_AGG_PF 4
#endif
constexpr int kAggPrefetchPasses = GKG_AGG_PF;                  // L2 prefetch distance, in passes of the warp

__device__ __forceinline__ uint32_t bf2_gt_mask(uint32_t a, uint32_t b) {
  return __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}

template <int K, int CPN, int SUB, bool ARG>
__global__ void __launch_bounds__(1024, 1)
mr_aggregate_fwd_bf16_warp_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_sb, int64_t x_sn,
                                  const __nv_bfloat16* __restrict__ y, int64_t y_sb, int64_t y_sn,
                                  const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ out,
                                  uint8_t* __restrict__ argmax, int B, int G, int N, int M, int D, int wide_store) {
  constexpr int CS = CPN * 8;                                   // channels per slice
  constexpr int NPW = 32 / CPN;                                 // nodes per warp pass
  constexpr int H = (K + 1) / 2;                                // id pairs (j, j + H) per list
  constexpr int NR = (H + SUB - 1) / SUB;                       // pair registers per lane
  constexpr bool SPLIT = CPN == 10;                             // main (8 chunks) + aux (2 chunks) key arrays
  extern __shared__ __align__(16) uint8_t agg_smem[];
  const int C = G * D, NSL = C / CS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // lanes past NPW * CPN shadow the last working lane (same addresses: no extra sectors, no divergence)
  const bool lane_on = lane < NPW * CPN;
  const int eff = lane_on ? lane : NPW * CPN - 1;
  int q, ch;                                                    // node within the pass, chunk within the slice
  if (SPLIT) {
    if (eff < 24) { q = eff >> 3; ch = eff & 7; } else { q = (eff - 24) >> 1; ch = 8 + ((eff - 24) & 1); }
  } else {
    q = eff / CPN; ch = eff - q * CPN;
  }
  const int slice = blockIdx.x % NSL;
  const int range = blockIdx.x / NSL, nranges = gridDim.x / NSL;
  if (range >= nranges) return;
  const int c0 = slice * CS, cc = c0 + ch * 8;
  const int g = cc / D;                                         // channel group of this lane's chunk
  const int s = ch % SUB, sub0 = ch - s;                        // chunks [sub0, sub0 + SUB) share a neighbour list
  auto src_lane = [&](int u) {                                  // lane that owns chunk sub0 + u of my node
    const int ci = sub0 + u;
    return SPLIT ? (ci < 8 ? 8 * q + ci : 16 + 2 * q + ci) : q * CPN + ci;
  };
  const long long R = (long long)B * N;
  const long long per = (R + nranges - 1) / nranges;
  long long r = (long long)range * per;
  const long long r_end = r + per < R ? r + per : R;
  const uint32_t keys_s = (uint32_t)__cvta_generic_to_shared(agg_smem);
  const uint32_t aux_s = keys_s + (uint32_t)M * 128;
  const uint32_t pitch = SPLIT ? (ch < 8 ? 128u : 32u) : (uint32_t)CS * 2;
  const uint32_t my_s = SPLIT ? (ch < 8 ? keys_s + ch * 16 : aux_s + (ch - 8) * 16) : keys_s + ch * 16;
  while (r < r_end) {
    const long long b = r / N;
    const int n_lo = (int)(r - b * N);
    const int n_hi = (int)((long long)n_lo + (r_end - r) < (long long)N ? n_lo + (r_end - r) : N);
    __syncthreads();                                            // gathers of the previous image are done
    {
      const __nv_bfloat16* src = y + b * y_sb + c0;
      const int pieces = M * CPN;
      for (int p = threadIdx.x; p < pieces; p += blockDim.x) {
        const int m = p / CPN, pc = p - m * CPN;
        const uint32_t dst = SPLIT ? (pc < 8 ? keys_s + (uint32_t)m * 128 + pc * 16 : aux_s + (uint32_t)m * 32 + (pc - 8) * 16)
                                   : keys_s + (uint32_t)p * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (long long)m * y_sn + pc * 8)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int32_t* ib = idx + ((b * G + g) * N) * (long long)K + s;
    const __nv_bfloat16* xb = x + b * x_sb + cc;
    const int npass = (n_hi - n_lo + NPW - 1) / NPW;
    // ids of pair t = s + i * SUB: t (later the low half) and t + H (high half); packed only when the
    // registers rotate, so that nothing waits on these loads inside the pass that issues them
    auto load_idx = [&](int node, uint32_t (&lo)[NR], uint32_t (&hi)[NR]) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int t = i * SUB + s;
        const int32_t* ip = ib + (long long)node * K + i * SUB;
        lo[i] = (node < n_hi && t < H) ? (uint32_t)__ldg(ip) : 0u;
        hi[i] = (node < n_hi && t < H && t + H < K) ? (uint32_t)__ldg(ip + H) : 0u;
      }
    };
    auto load_x = [&](int node) {
      return node < n_hi ? *reinterpret_cast<const uint4*>(xb + (long long)node * x_sn) : make_uint4(0, 0, 0, 0);
    };
    auto prefetch_l2 = [&](int node) {                           // DRAM -> L2 a few passes ahead (no registers held)
      if (node < n_hi) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + (long long)node * x_sn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ib + (long long)node * K));
      }
    };
    int pass = warp;
    uint32_t ir[NR], irn[NR], irh[NR];
    uint4 xv;
    load_idx(n_lo + pass * NPW + q, irn, irh);
    xv = load_x(n_lo + pass * NPW + q);
#pragma unroll
    for (int d = 1; d < kAggPrefetchPasses; ++d) prefetch_l2(n_lo + (pass + d * nwarps) * NPW + q);
#pragma unroll
    for (int i = 0; i < NR; ++i) ir[i] = irn[i] | (irh[i] << 16);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                            // keys have landed
    for (; pass < npass; pass += nwarps) {
      const int node = n_lo + pass * NPW + q;
      load_idx(node + nwarps * NPW, irn, irh);
      const uint4 xn = load_x(node + nwarps * NPW);
      prefetch_l2(node + kAggPrefetchPasses * nwarps * NPW);
      uint32_t pv[H];
#pragma unroll
      for (int t = 0; t < H; ++t) pv[t] = __shfl_sync(0xffffffffu, ir[t / SUB], src_lane(t % SUB));
      uint32_t b0, b1, b2, b3, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) {                              // in list order: the smallest j wins ties
        const uint32_t m = j < H ? (pv[j] & 0xffffu) : (pv[j - H] >> 16);
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(my_s + m * pitch));
        if (j == 0) {
          b0 = v0; b1 = v1; b2 = v2; b3 = v3;
        } else {
          if (ARG) {
            const uint32_t jj = (uint32_t)j * 0x00010001u;
            uint32_t t;
            t = bf2_gt_mask(v0, b0); a0 = (a0 & ~t) | (jj & t);
            t = bf2_gt_mask(v1, b1); a1 = (a1 & ~t) | (jj & t);
            t = bf2_gt_mask(v2, b2); a2 = (a2 & ~t) | (jj & t);
            t = bf2_gt_mask(v3, b3); a3 = (a3 & ~t) | (jj & t);
          }
          b0 = bf2_max(b0, v0); b1 = bf2_max(b1, v1); b2 = bf2_max(b2, v2); b3 = bf2_max(b3, v3);
        }
      }
      if (lane_on && node < n_hi) {
        uint4 o0, o1;
        bf2_emit(xv.x, b0, o0.x, o0.y); bf2_emit(xv.y, b1, o0.z, o0.w);
        bf2_emit(xv.z, b2, o1.x, o1.y); bf2_emit(xv.w, b3, o1.z, o1.w);
        const long long bn = b * N + node;
        __nv_bfloat16* op = out + (bn * C + cc) * 2;
        if (wide_store) {
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(o0.x), "r"(o0.y),
                       "r"(o0.z), "r"(o0.w), "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w) : "memory");
        } else {
          reinterpret_cast<uint4*>(op)[0] = o0;
          reinterpret_cast<uint4*>(op)[1] = o1;
        }
        if (ARG) {
          uint2 a;
          a.x = __byte_perm(a0, a1, 0x6420);
          a.y = __byte_perm(a2, a3, 0x6420);
          *reinterpret_cast<uint2*>(argmax + bn * C + cc) = a;
        }
      }
      xv = xn;
#pragma unroll
      for (int i = 0; i < NR; ++i) ir[i] = irn[i] | (irh[i] << 16);
    }
    r += n_hi - n_lo;
  }
}

// Slice geometry of the warp-autonomous kernel: CS in {80, 40} channels, a slice either holds whole
// channel groups or lies inside one, neighbour lists shared by SUB in {5, 10} lanes.
struct AggWarpPlan { int cpn, sub, ctas_per_sm, threads; size_t smem; };
static bool agg_warp_plan(int N, int M, int D, int C, AggWarpPlan* p) {
  if (N < 1024 || N < M || M > 65535) return false;  // label heads: few queries, nothing to amortise; ids travel as 16 bits
  for (int cs = 80; cs >= 40; cs -= 40) {
    if (C % cs) continue;
    int sub;
    if (D % cs == 0) sub = cs / 8; else if (cs % D == 0 && D % 8 == 0) sub = D / 8; else continue;
    if (sub != 5 && sub != 10) continue;
    const size_t smem = (size_t)M * cs * 2;
    if (smem > kAggWarpSmemMax) continue;
    p->cpn = cs / 8; p->sub = sub; p->smem = smem;
    p->ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
    p->threads = p->ctas_per_sm == 2 ? 512 : 1024;
    return true;
  }
  return false;
}

// Channel-slice width for the shared-memory path: the widest multiple of 8 that divides D (a slice
// stays inside one channel group) whose keys fit; 0 = use the L2 gather kernel.
static int agg_smem_slice(int N, int M, int D) {
  if (N < 1024 || N < M) return 0;                  // label heads: few queries, nothing to amortise
  for (int cs = D; cs >= 8; --cs) {
    if (D % cs || cs % 8) continue;
    if ((size_t)M * cs * 2 <= kAggSmemKeysMax) return cs;
  }
  return 0;
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
#define LAUNCH_BF(KK, AA)                                                                       \
  mr_aggregate_fwd_bf16_kernel<KK, AA><<<grid, 256, 0, stream>>>(                                 \
      static_cast<const __nv_bfloat16*>(x), x_sb, x_sn, static_cast<const __nv_bfloat16*>(y), y_sb, \
      y_sn, idx, static_cast<__nv_bfloat16*>(out), argmax, G, N, D, chunks)
  if (dtype == GKG_F32) {
    if (vec == 4) LAUNCH(float, 4); else if (vec == 2) LAUNCH(float, 2); else LAUNCH(float, 1);
  } else if (vec == 8 && (k == 9 || k == 18)) {
    const int cs = agg_smem_slice(N, M, D);
    AggWarpPlan wp;
    if (agg_warp_plan(N, M, D, C, &wp)) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int nsl = C / (wp.cpn * 8), npw = 32 / wp.cpn;
      int nranges = wp.ctas_per_sm * sms / nsl;
      if (nranges < 1) nranges = 1;
      const long long passes = ((long long)B * N + npw - 1) / npw;
      if (nranges > passes) nranges = (int)passes;
      const int wgrid = nranges * nsl;
      const int wide = ((uintptr_t)out % 32) == 0;
#define LAUNCH_W(KK, CP, SB, AA)                                                                    \
  do {                                                                                               \
    auto kern = mr_aggregate_fwd_bf16_warp_kernel<KK, CP, SB, AA>;                                   \
    static std::atomic<uint64_t> configured{0};                                                     \
    cudaError_t e = cudaSuccess;                                                                    \
    configure_once_per_device(configured, [&] {                                                     \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAggWarpSmemMax); }); \
    if (e != cudaSuccess) { set_error("mr_aggregate_fwd: smem attribute: %s", cudaGetErrorString(e)); return GKG_ECUDA; } \
    kern<<<wgrid, wp.threads, wp.smem, stream>>>(static_cast<const __nv_bfloat16*>(x), x_sb, x_sn,   \
        static_cast<const __nv_bfloat16*>(y), y_sb, y_sn, idx, static_cast<__nv_bfloat16*>(out), argmax, \
        B, G, N, M, D, wide);                                                                        \
  } while (0)
#define LAUNCH_WK(KK, AA)                                                       \
  do {                                                                           \
    if (wp.cpn == 10 && wp.sub == 5) LAUNCH_W(KK, 10, 5, AA);                    \
    else if (wp.cpn == 10) LAUNCH_W(KK, 10, 10, AA);                             \
    else LAUNCH_W(KK, 5, 5, AA);                                                 \
  } while (0)
      if (k == 9) { if (argmax) LAUNCH_WK(9, true); else LAUNCH_WK(9, false); }
      else { if (argmax) LAUNCH_WK(18, true); else LAUNCH_WK(18, false); }
#undef LAUNCH_WK
#undef LAUNCH_W
    } else if (cs > 0) {
      const int nsl = C / cs, npp = kAggSmemThreads / (cs / 8);
      const size_t smem = (size_t)M * cs * 2 + 2 * (size_t)npp * k * 4;
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int nranges = 2 * sms / nsl;
      if (nranges < 1) nranges = 1;
      if ((long long)nranges * npp > (long long)B * N) nranges = (int)(((long long)B * N + npp - 1) / npp);
      const int sgrid = nranges * nsl;
#define LAUNCH_SM(KK, AA)                                                                          \
  do {                                                                                               \
    auto kern = mr_aggregate_fwd_bf16_smem_kernel<KK, AA>;                                           \
    static std::atomic<uint64_t> configured{0};                                                     \
    cudaError_t e = cudaSuccess;                                                                    \
    configure_once_per_device(configured, [&] {                                                     \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kAggSmemKeysMax + 16 * 1024)); }); \
    if (e != cudaSuccess) { set_error("mr_aggregate_fwd: smem attribute: %s", cudaGetErrorString(e)); return GKG_ECUDA; } \
    kern<<<sgrid, kAggSmemThreads, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), x_sb, x_sn,  \
        static_cast<const __nv_bfloat16*>(y), y_sb, y_sn, idx, static_cast<__nv_bfloat16*>(out), argmax, \
        B, G, N, M, D, cs);                                                                          \
  } while (0)
      if (k == 9) { if (argmax) LAUNCH_SM(9, true); else LAUNCH_SM(9, false); }
      else { if (argmax) LAUNCH_SM(18, true); else LAUNCH_SM(18, false); }
#undef LAUNCH_SM
    } else if (k == 9) { if (argmax) LAUNCH_BF(9, true); else LAUNCH_BF(9, false); }
    else { if (argmax) LAUNCH_BF(18, true); else LAUNCH_BF(18, false); }
  } else {
    if (vec == 8) LAUNCH(__nv_bfloat16, 8); else if (vec == 4) LAUNCH(__nv_bfloat16, 4);
    else if (vec == 2) LAUNCH(__nv_bfloat16, 2); else LAUNCH(__nv_bfloat16, 1);
  }
#undef LAUNCH
#undef LAUNCH_BF
  GKG_CHECK_LAUNCH("mr_aggregate_fwd");
  return GKG_OK;
}

static int mr_aggregate_bwd_impl(const void* grad_out, const int32_t* idx, const uint8_t* argmax, void* grad_x,
                                 float* grad_y_accum, const float* scale, int B, int G, int N, int M, int D, int k,
                                 int dtype, cudaStream_t stream) {
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
  do {                                                                                         \
    if (scale != nullptr)                                                                      \
      mr_aggregate_bwd_kernel<T, V, true><<<grid, 256, 0, stream>>>(                           \
          static_cast<const T*>(grad_out), idx, argmax, static_cast<T*>(grad_x), grad_y_accum, scale, G, N, M, D, k, chunks); \
    else                                                                                       \
      mr_aggregate_bwd_kernel<T, V, false><<<grid, 256, 0, stream>>>(                          \
          static_cast<const T*>(grad_out), idx, argmax, static_cast<T*>(grad_x), grad_y_accum, nullptr, G, N, M, D, k, chunks); \
  } while (0)
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

extern "C" int gkg_mr_aggregate_bwd(const void* grad_out, const int32_t* idx, const uint8_t* argmax,
                                    void* grad_x, float* grad_y_accum, int B, int G, int N, int M,
                                    int D, int k, int dtype, gkg_stream_t stream) {
  return mr_aggregate_bwd_impl(grad_out, idx, argmax, grad_x, grad_y_accum, nullptr, B, G, N, M, D, k, dtype,
                               static_cast<cudaStream_t>(stream));
}

// Deterministic form: grad_y is accumulated in 64-bit fixed point (see the kernel) and converted afterwards.
extern "C" int gkg_mr_aggregate_bwd_det(const void* grad_out, const int32_t* idx, const uint8_t* argmax,
                                        void* grad_x, long long* grad_y_fixed, const float* scale, int B, int G, int N,
                                        int M, int D, int k, int dtype, gkg_stream_t stream) {
  GKG_CHECK_ARG(scale != nullptr, "mr_aggregate_bwd_det: null scale");
  return mr_aggregate_bwd_impl(grad_out, idx, argmax, grad_x, reinterpret_cast<float*>(grad_y_fixed), scale, B, G, N, M,
                               D, k, dtype, static_cast<cudaStream_t>(stream));
}

namespace gkg {
__global__ void fixed_to_float_kernel(const long long* __restrict__ in, const float* __restrict__ scale,
                                      float* __restrict__ out, long long n) {
  const float inv = 1.f / __ldg(scale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)((double)in[i] * (double)inv);
}
}  // namespace gkg

extern "C" int gkg_fixed_to_float(const long long* in, const float* scale, float* out, long long n, gkg_stream_t stream) {
  GKG_CHECK_ARG(n >= 0 && (n == 0 || (in && scale && out)), "fixed_to_float: bad arguments");
  if (n == 0) return GKG_OK;
  gkg::fixed_to_float_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, scale, out, n);
  GKG_CHECK_LAUNCH("fixed_to_float_kernel");
  return GKG_OK;
}

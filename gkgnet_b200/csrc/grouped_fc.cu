// Grouped 1x1 FC of the max-relative graph convolution on the Blackwell tensor cores, with the
// normalisation and activation fused into the epilogue (inference form).
//
// Reference: MRConv2d.nn = BasicConv([2C, 2C]) = Conv2d(2C, 2C, 1, groups=4, bias) -> norm -> act
// (torch_nn.py:57-81, used at torch_vertex.py:45,61).  In eval mode the batch norm is an affine map per
// output channel, so the whole stack is
//     out[r, o] = act( sum_i (scale[o] * W[o, i]) * in[r, q*CG + i] + shift[o] ),   q = o / CG, CG = 2C / 4
// with scale = gamma / sqrt(var + eps) folded into the weights and shift = (bias - mean) * scale + beta by the caller.
// One pass over the (rows, 2C) activation instead of the three of conv / norm / act.
//
// Kernel: a CTA owns tiles of 128 rows.  All threads copy the tile into shared memory in the UMMA K-major
// core-matrix order (8 rows x 16 bytes contiguous; one A sub-tile of KP = ceil16(CG) columns per conv
// group, zero padded), one thread issues tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate; M = 128,
// N = NP = ceil16(CG), K = 16 per instruction) for the four groups into 4 * NP TMEM columns, the warps read
// their accumulator rows back with tcgen05.ld (thread == row, one conv group per warp), apply scale / shift / activation and store
// bf16.  The weights sit in shared memory for the CTA's lifetime in the same core-matrix order (built by
// the caller, see gkgnet_b200/ops.py:grouped_fc_weights).  Memory bound: 2 * rows * 2C * 2 bytes.
// Three generations of the narrow kernel live here: grouped_fc_kernel (synchronous, small launches), grouped_fc_pipe_kernel
// (cp.async double buffering; fallback when the tensor map cannot be encoded) and grouped_fc_tma_kernel (TMA copy-in, control
// warp, bulk-store copy-out: >= 2 tiles per SM); grouped_fc_wide_kernel covers CG > 96, grouped_fc_wgrad_kernel the weight gradient.
#include "knn_tc.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace gkg {
namespace fc {

constexpr int BM = 128;
constexpr int THREADS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // cute::UMMA::SmemDescriptor, no swizzle: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned long long spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1ull << 26)) __trap();        // a protocol bug must abort, never hang the GPU
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 resolution): one reciprocal, one
// exponential and a degree-5 polynomial instead of libdevice's erff (half the instructions of an epilogue
// that evaluates 2C of them per row)
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.f - p * t * __expf(-ax * ax);
  return copysignf(e, x);
}
// nn.GELU() (erf form) = v * Phi(v) with the same 7.1.26 polynomial arranged for the fewest instructions:
//   Phi(-|v|) = 0.5 erfc(|v| / sqrt 2) = (0.5 (a1 t + .. + a5 t^5)) * 2^(-(|v| sqrt(log2(e) / 2))^2),  t = 1 / (1 + (p / sqrt 2) |v|)
// 13 FMA-pipe operations + a compare + 2 MUFU per element (the 0.5 v (1 + erf(v / sqrt 2)) form took ~20: the epilogue of the
// fused FC is issue bound).  |error of Phi| <= 7.5e-8.
__device__ __forceinline__ float gelu_erf(float v) {
  const float a = fabsf(v);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.231641888f, a, 1.f)));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float sq = a * 0.849321800f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-sq * sq));
  const float h = p * t * e;
  return v * (v >= 0.f ? 1.f - h : h);
}
template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return gelu_erf(v);
  return v;
}

struct Params {
  const __nv_bfloat16* in;       // (rows, C2) contiguous
  __nv_bfloat16* out;            // (rows, C2) contiguous
  const __nv_bfloat16* w_op;     // 4 groups x [NP/8][KP/8][8][8] core matrices (K-major)
  const float* shift;            // (C2); the per-channel scale is folded into w_op by the caller
  long long rows;
  int C2, CG, KP, NP, act;
  int stages;                    // TMA kernel: depth of the A-tile ring
  int stage_out;                 // TMA kernel: output tiles leave through the consumed stage + one bulk store each
};

// CGT: channels per conv group at compile time (the two wide layers), 0 = run-time value
template <int CGT, int ACT>
__global__ void __launch_bounds__(THREADS, 2) grouped_fc_kernel(const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int CG = CGT > 0 ? CGT : prm.CG;
  const int KP = (CG + 15) / 16 * 16, NP = KP, C2 = 4 * CG;
  const uint32_t a_group_bytes = (uint32_t)BM * KP * 2;
  const uint32_t b_group_bytes = (uint32_t)NP * KP * 2;
  uint8_t* sA = smem;                                   // 4 x [BM/8][KP/8][8][8]
  uint8_t* sB = sA + 4 * a_group_bytes;                 // 4 x [NP/8][KP/8][8][8]
  float* s_shift = reinterpret_cast<float*>(sB + 4 * b_group_bytes);   // (the scale is folded into the weights)
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_shift + C2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = 4 * NP <= 256 ? 256u : 512u;   // power of two; two CTAs share an SM when 256 suffice

  // ---- one-time setup: weights, affine vectors, barrier, TMEM
  for (uint32_t i = threadIdx.x; i < 4 * b_group_bytes / 16; i += THREADS)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(prm.w_op) + i);
  for (int i = threadIdx.x; i < C2; i += THREADS) s_shift[i] = prm.shift[i];
  // zero the K padding of the A sub-tiles once (columns CG..KP never change)
  for (uint32_t i = threadIdx.x; i < 4 * a_group_bytes / 16; i += THREADS)
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // kind::f16 instruction descriptor: D = f32 (bit 4), A = B = bf16 (1 at bits 7 and 10), K-major, N>>3 at 17, M>>4 at 24
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t sbo = (uint32_t)(KP >> 3) * 128u;      // bytes between 8-row groups
  const int chunks_per_row = C2 >> 3;                   // 16-byte pieces of an input row
  const int cg_chunks = CG >> 3;                        // ... per conv group
  const long long tiles = (prm.rows + BM - 1) / BM;
  uint32_t phase = 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r0 = tile * BM;
    // ---- rows -> shared memory, core-matrix order: piece (r, c8) of group q lands at
    //      q*a_group + ((r/8)*(KP/8) + c8)*128 + (r%8)*16
#pragma unroll 5
    for (int i = threadIdx.x; i < BM * chunks_per_row; i += THREADS) {
      const int r = i / chunks_per_row, cc = i - r * chunks_per_row;
      const int q = cc / cg_chunks, c8 = cc - q * cg_chunks;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r0 + r < prm.rows) v = __ldg(reinterpret_cast<const uint4*>(prm.in + (r0 + r) * C2) + cc);
      *reinterpret_cast<uint4*>(sA + q * a_group_bytes + ((uint32_t)((r >> 3) * (KP >> 3) + c8) << 7) + ((r & 7) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int q = 0; q < 4; ++q) {
        const uint32_t a_addr = smem_u32(sA + q * a_group_bytes), b_addr = smem_u32(sB + q * b_group_bytes);
        for (int ks = 0; ks < (KP >> 4); ++ks) {          // one K = 16 step = two core matrices = 256 bytes
          const uint64_t ad = make_desc(a_addr + ks * 256, 128, sbo), bd = make_desc(b_addr + ks * 256, 128, sbo);
          const uint32_t acc = ks != 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem_base + (uint32_t)(q * NP)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(smem_u32(bar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., conv group w/4
    {
      const int row = (warp & 3) * 32 + lane;
      const bool row_ok = r0 + row < prm.rows;
      __nv_bfloat16* orow = prm.out + (r0 + row) * C2;
      {
        const int q = warp >> 2;
        for (int c0 = 0; c0 < CG; c0 += 16) {             // CG is a multiple of 8: the last piece may be half valid
          uint32_t acc[16];
          tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(q * NP + c0), acc);
          __align__(16) __nv_bfloat162 o[8];
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {             // CG is a multiple of 8: groups of 4 never straddle the end
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (c0 + j4 < CG) {
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + q * CG + c0 + j4);
              v[0] = activate<ACT>(__uint_as_float(acc[j4]) + sh.x);
              v[1] = activate<ACT>(__uint_as_float(acc[j4 + 1]) + sh.y);
              v[2] = activate<ACT>(__uint_as_float(acc[j4 + 2]) + sh.z);
              v[3] = activate<ACT>(__uint_as_float(acc[j4 + 3]) + sh.w);
            }
            o[j4 >> 1] = __floats2bfloat162_rn(v[0], v[1]);
            o[(j4 >> 1) + 1] = __floats2bfloat162_rn(v[2], v[3]);
          }
          if (row_ok) {
            *reinterpret_cast<uint4*>(orow + q * CG + c0) = *reinterpret_cast<const uint4*>(o);
            if (c0 + 8 < CG) *reinterpret_cast<uint4*>(orow + q * CG + c0 + 8) = *reinterpret_cast<const uint4*>(o + 4);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                      // TMEM and the A tile are free again
  }

  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}


// ------------------------------------------------------------------------------------
// narrow groups, software pipelined (round 2)
// ------------------------------------------------------------------------------------
// Same arithmetic and shared-memory formats as grouped_fc_kernel, ONE persistent CTA per SM with two A buffers and (when
// eight accumulators fit: CG <= 64) two sets of TMEM accumulators.  Iteration i:
//     wait for the cp.async copies of tile i+1  ->  barrier  ->  issue the MMAs of tile i+1  ->  wait for the MMAs of
//     tile i  ->  start the copies of tile i+2 into the buffer tile i just left  ->  epilogue of tile i
// so the global loads of a tile have a whole iteration to land and the MMA + commit round trip of the next tile runs
// under the epilogue of this one.  With one accumulator set (CG = 80) the MMAs of tile i+1 are issued after the
// epilogue of tile i instead (its loads still overlap).  The synchronous kernel above spent half its time waiting
// for the copy-in and the MMA round trip (0.19 ms at the bench shape against 65 us of HBM time).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int CGT, int ACT>
__global__ void __launch_bounds__(THREADS, 1) grouped_fc_pipe_kernel(const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int CG = CGT > 0 ? CGT : prm.CG;
  const int KP = (CG + 15) / 16 * 16, NP = KP, C2 = 4 * CG;
  const bool two_acc = 8 * NP <= 512;
  const uint32_t a_group_bytes = (uint32_t)BM * KP * 2;
  const uint32_t b_group_bytes = (uint32_t)NP * KP * 2;
  uint8_t* sA = smem;                                   // 2 x 4 x [BM/8][KP/8][8][8]
  uint8_t* sB = sA + 8 * a_group_bytes;                 // 4 x [NP/8][KP/8][8][8]
  float* s_shift = reinterpret_cast<float*>(sB + 4 * b_group_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_shift + C2);      // [2]: MMAs of the tile in accumulator set 0 / 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (uint32_t i = threadIdx.x; i < 4 * b_group_bytes / 16; i += THREADS)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(prm.w_op) + i);
  for (int i = threadIdx.x; i < C2; i += THREADS) s_shift[i] = prm.shift[i];
  for (uint32_t i = threadIdx.x; i < 8 * a_group_bytes / 16; i += THREADS)      // K padding of both buffers stays zero
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init(smem_u32(bar + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t sbo = (uint32_t)(KP >> 3) * 128u;
  const int chunks_per_row = C2 >> 3, cg_chunks = CG >> 3;
  const long long tiles = (prm.rows + BM - 1) / BM;
  const long long first = blockIdx.x, stride = gridDim.x;
  const long long mine = first < tiles ? (tiles - first + stride - 1) / stride : 0;     // tiles of this CTA

  // copy-in: thread t owns the 16-byte piece cc = t % chunks_per_row of the rows rl, rl + rpp, ... (rpp rows per pass):
  // the piece's conv group and position inside the core-matrix layout are fixed per thread
  const int rpp = THREADS / chunks_per_row;
  const int ld_rl = threadIdx.x / chunks_per_row, ld_cc = threadIdx.x - ld_rl * chunks_per_row;
  const bool ld_on = ld_rl < rpp;
  const uint32_t ld_col = (uint32_t)(ld_cc / cg_chunks) * a_group_bytes + ((uint32_t)(ld_cc % cg_chunks) << 7);
  auto load_tile = [&](long long j) {                   // j-th tile of this CTA -> A buffer j & 1 (asynchronous)
    const long long r0 = (first + j * stride) * BM;
    const uint32_t base = smem_u32(sA) + (uint32_t)(j & 1) * 4 * a_group_bytes + ld_col;
    if (ld_on) {
      const __nv_bfloat16* src = prm.in + (r0 + ld_rl) * C2 + ld_cc * 8;
      const long long step = (long long)rpp * C2;
      if (r0 + BM <= prm.rows) {                        // every tile but (possibly) the last: no row checks
#pragma unroll 2
        for (int r = ld_rl; r < BM; r += rpp, src += step)
          cp_async16(base + ((uint32_t)((r >> 3) * (KP >> 3)) << 7) + ((r & 7) << 4), src, 16u);
      } else {
        for (int r = ld_rl; r < BM; r += rpp, src += step) {
          const bool ok = r0 + r < prm.rows;
          cp_async16(base + ((uint32_t)((r >> 3) * (KP >> 3)) << 7) + ((r & 7) << 4), ok ? src : prm.in, ok ? 16u : 0u);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto issue_mma = [&](long long j) {                   // one thread; tile j: A buffer j & 1 -> accumulator set
    const uint32_t acc_set = two_acc ? (uint32_t)(j & 1) : 0u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int q = 0; q < 4; ++q) {
      const uint32_t a_addr = smem_u32(sA) + (uint32_t)(j & 1) * 4 * a_group_bytes + q * a_group_bytes;
      const uint32_t b_addr = smem_u32(sB + q * b_group_bytes);
      for (int ks = 0; ks < (KP >> 4); ++ks) {
        const uint64_t ad = make_desc(a_addr + ks * 256, 128, sbo), bd = make_desc(b_addr + ks * 256, 128, sbo);
        const uint32_t acc = ks != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base + acc_set * 4 * NP + (uint32_t)(q * NP)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + acc_set)) : "memory");
  };
  auto epilogue = [&](long long j) {
    const long long r0 = (first + j * stride) * BM;
    const uint32_t acc_set = two_acc ? (uint32_t)(j & 1) : 0u;
    const int row = (warp & 3) * 32 + lane;
    const bool row_ok = r0 + row < prm.rows;
    __nv_bfloat16* orow = prm.out + (r0 + row) * C2;
    const int q = warp >> 2;
    for (int c0 = 0; c0 < CG; c0 += 16) {
      uint32_t acc[16];
      tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc_set * 4 * NP + (uint32_t)(q * NP + c0), acc);
      __align__(16) __nv_bfloat162 o[8];
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c0 + j4 < CG) {
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + q * CG + c0 + j4);
          v[0] = activate<ACT>(__uint_as_float(acc[j4]) + sh.x);
          v[1] = activate<ACT>(__uint_as_float(acc[j4 + 1]) + sh.y);
          v[2] = activate<ACT>(__uint_as_float(acc[j4 + 2]) + sh.z);
          v[3] = activate<ACT>(__uint_as_float(acc[j4 + 3]) + sh.w);
        }
        o[j4 >> 1] = __floats2bfloat162_rn(v[0], v[1]);
        o[(j4 >> 1) + 1] = __floats2bfloat162_rn(v[2], v[3]);
      }
      if (row_ok) {
        *reinterpret_cast<uint4*>(orow + q * CG + c0) = *reinterpret_cast<const uint4*>(o);
        if (c0 + 8 < CG) *reinterpret_cast<uint4*>(orow + q * CG + c0 + 8) = *reinterpret_cast<const uint4*>(o + 4);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  uint32_t ph[2] = {0, 0};
  auto wait_mma = [&](long long j) {
    const uint32_t acc_set = two_acc ? (uint32_t)(j & 1) : 0u;
    mbar_wait(smem_u32(bar + acc_set), ph[acc_set]);
    ph[acc_set] ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // every thread's copies of the OLDER tile done and visible to the MMA (newer: one more group may stay in flight)
  auto loads_landed = [&](bool newer_in_flight) {
    if (newer_in_flight) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  };

  if (mine > 0) {
    load_tile(0);
    loads_landed(false);
    if (threadIdx.x == 0) issue_mma(0);
    if (mine > 1) load_tile(1);
    for (long long j = 0; j < mine; ++j) {
      if (two_acc) {
        if (j + 1 < mine) {
          loads_landed(false);                           // tile j+1 in its buffer; every warp is past the epilogue of j-1
          if (threadIdx.x == 0) issue_mma(j + 1);
        }
        wait_mma(j);                                     // buffer j & 1 has been read: refill it
        if (j + 2 < mine) load_tile(j + 2);
        epilogue(j);
      } else {
        wait_mma(j);
        if (j + 2 < mine) load_tile(j + 2);              // (buffer j & 1 is free; tile j+1 was started an iteration ago)
        epilogue(j);
        if (j + 1 < mine) {
          loads_landed(j + 2 < mine);                    // also: every warp has drained the accumulators of tile j
          if (threadIdx.x == 0) issue_mma(j + 1);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// ------------------------------------------------------------------------------------
// the same pipeline with the copy-in on the TMA: a 3-D tensor map over the input, (8 channels, rows, C2 / 8) with box
// (8, BM, C2 / 8), lands a tile as [C2 / 8][BM][16 bytes].  That IS a no-swizzle K-major operand: the 8 rows of a core
// matrix are 128 contiguous bytes, 8-row groups are 128 bytes apart (SBO), K-adjacent core matrices BM * 16 bytes (LBO).
// One instruction per tile instead of BM * C2 / 8 per-thread cp.async: the per-thread 16-byte path saturates near
// 4 TB/s chip-wide.  The K padding of a group (CG = 40: 40 -> 48) reads the first 8 channels of the NEXT
// group against zero weight rows (finite x 0; the last group reads the 2 KB zero tail behind the ring... of the last
// stage only -- the other stages are followed by the next stage's tile, also finite), rows past the end are zero-filled
// by the TMA.  NST stages: tiles j+1 .. j+NST-1 are in flight during the epilogue of tile j.
template <int CGT, int ACT>
__global__ void __launch_bounds__(THREADS + 32, 1) grouped_fc_tma_kernel(const Params prm, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int CG = CGT > 0 ? CGT : prm.CG;
  const int KP = (CG + 15) / 16 * 16, NP = KP, C2 = 4 * CG;
  const bool two_acc = 8 * NP <= 512;
  const int NST = prm.stages;
  const uint32_t tile_bytes = (uint32_t)BM * C2 * 2;
  const uint32_t b_group_bytes = (uint32_t)NP * KP * 2;
  uint8_t* sA = smem;                                   // NST x [C2/8][BM][8] + 2 KB of zeros
  uint8_t* sB = sA + (uint32_t)NST * tile_bytes + 2048; // 4 x [NP/8][KP/8][8][8]
  float* s_shift = reinterpret_cast<float*>(sB + 4 * b_group_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_shift + C2);      // [2]: MMAs of the tile in accumulator set 0 / 1
  uint64_t* full = bar + 2;                                       // [NST]: tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warp 16 (one thread) is the control thread: TMA loads, MMA issue, bulk stores; warps 0-15 only wait for the MMAs and run
  // the epilogue.  stage_out: the output tile is written into the A stage the MMA has just consumed (same size, row-major
  // like the output itself) and leaves as ONE bulk store; the stage is refilled once the store has read it.
  const bool ctl = threadIdx.x == THREADS;
  const bool stage_out = two_acc && NST >= 3 && prm.stage_out != 0;

  for (uint32_t i = threadIdx.x; i < 4 * b_group_bytes / 16; i += THREADS + 32)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(prm.w_op) + i);
  for (int i = threadIdx.x; i < C2; i += THREADS + 32) s_shift[i] = prm.shift[i];
  // every stage starts finite (the K padding of the last group reads the head of the NEXT stage, loaded or not) + the tail
  for (uint32_t i = threadIdx.x; i < ((uint32_t)NST * tile_bytes + 2048) / 16; i += THREADS + 32)
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init(smem_u32(bar + 1), 1);
    for (int i = 0; i < NST; ++i) mbar_init(smem_u32(full + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // zero tail / weights written by this proxy, read by the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t sbo_b = (uint32_t)(KP >> 3) * 128u;
  const int cg_chunks = CG >> 3;
  const long long tiles = (prm.rows + BM - 1) / BM;
  const long long first = blockIdx.x, stride = gridDim.x;
  const int mine = first < tiles ? (int)((tiles - first + stride - 1) / stride) : 0;     // tiles of this CTA (32-bit loop
  // counters: `j % NST` on a long long is a 64-bit division routine, and it sat in the epilogue's inner loop)

  auto load_tile = [&](int j) {                   // one thread; j-th tile of this CTA -> stage j % NST
    const int s = j % NST;
    const int r0 = (int)((first + j * stride) * BM);
    const uint32_t fb = smem_u32(full + s);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(tile_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(sA) + (uint32_t)s * tile_bytes), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(r0), "r"(0),
                   "r"(fb) : "memory");
  };
  auto issue_mma = [&](int j) {                   // one thread; tile j: stage j % NST -> accumulator set
    const uint32_t acc_set = two_acc ? (uint32_t)(j & 1) : 0u;
    const int sj = j % NST;
    mbar_wait(smem_u32(full + sj), (uint32_t)((j / NST) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_tile = smem_u32(sA) + (uint32_t)sj * tile_bytes;
    for (int q = 0; q < 4; ++q) {
      const uint32_t a_addr = a_tile + (uint32_t)(q * cg_chunks) * (BM * 16);
      const uint32_t b_addr = smem_u32(sB + q * b_group_bytes);
      for (int ks = 0; ks < (KP >> 4); ++ks) {
        const uint64_t ad = make_desc(a_addr + ks * 2 * (BM * 16), BM * 16, 128), bd = make_desc(b_addr + ks * 256, 128, sbo_b);
        const uint32_t acc = ks != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base + acc_set * 4 * NP + (uint32_t)(q * NP)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + acc_set)) : "memory");
  };
  auto epilogue = [&](int j) {
    const long long r0 = (first + j * stride) * BM;
    const uint32_t acc_set = two_acc ? (uint32_t)(j & 1) : 0u;
    const int row = (warp & 3) * 32 + lane;
    const bool row_ok = r0 + row < prm.rows;
    __nv_bfloat16* orow = prm.out + (r0 + row) * C2;
    const int q = warp >> 2;
    uint8_t* srow = sA + (uint32_t)(j % NST) * tile_bytes + ((uint32_t)row * C2 + q * CG) * 2;   // stage_out
    for (int c0 = 0; c0 < CG; c0 += 16) {
      uint32_t acc[16];
      tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc_set * 4 * NP + (uint32_t)(q * NP + c0), acc);
      __align__(16) __nv_bfloat162 o[8];
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c0 + j4 < CG) {
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + q * CG + c0 + j4);
          v[0] = activate<ACT>(__uint_as_float(acc[j4]) + sh.x);
          v[1] = activate<ACT>(__uint_as_float(acc[j4 + 1]) + sh.y);
          v[2] = activate<ACT>(__uint_as_float(acc[j4 + 2]) + sh.z);
          v[3] = activate<ACT>(__uint_as_float(acc[j4 + 3]) + sh.w);
        }
        o[j4 >> 1] = __floats2bfloat162_rn(v[0], v[1]);
        o[(j4 >> 1) + 1] = __floats2bfloat162_rn(v[2], v[3]);
      }
      if (stage_out) {
        uint4* dst = reinterpret_cast<uint4*>(srow + c0 * 2);
        dst[0] = *reinterpret_cast<const uint4*>(o);
        if (c0 + 8 < CG) dst[1] = *reinterpret_cast<const uint4*>(o + 4);
      } else if (row_ok) {
        *reinterpret_cast<uint4*>(orow + q * CG + c0) = *reinterpret_cast<const uint4*>(o);
        if (c0 + 8 < CG) *reinterpret_cast<uint4*>(orow + q * CG + c0 + 8) = *reinterpret_cast<const uint4*>(o + 4);
      }
    }
    if (stage_out) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the bulk store
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };
  auto store_tile = [&](int j) {                  // control thread, after a barrier behind epilogue(j)
    const long long r0 = (first + j * stride) * BM;
    const long long nrows = prm.rows - r0 < BM ? prm.rows - r0 : BM;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(prm.out + r0 * C2), "r"(smem_u32(sA) + (uint32_t)(j % NST) * tile_bytes), "r"((uint32_t)(nrows * C2 * 2))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  };
  uint32_t ph[2] = {0, 0};
  auto wait_mma = [&](int j) {
    const uint32_t acc_set = two_acc ? (uint32_t)(j & 1) : 0u;
    mbar_wait(smem_u32(bar + acc_set), ph[acc_set]);
    ph[acc_set] ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };

  if (mine > 0) {
    if (ctl) {
      for (int t = 0; t < mine && t < NST; ++t) load_tile(t);
      issue_mma(0);
    }
    for (int j = 0; j < mine; ++j) {
      if (two_acc) {
        if (stage_out || j + 1 < mine) __syncthreads();  // every warp is past the epilogue of j-1: its accumulators are free
        if (ctl) {
          if (j + 1 < mine) issue_mma(j + 1);
          if (stage_out && j >= 1) {
            store_tile(j - 1);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stage has been read: refill it
            if (j - 1 + NST < mine) load_tile(j - 1 + NST);
          }
        }
        if (!stage_out || !ctl) wait_mma(j);             // stage j % NST has been read by the MMAs
        if (!stage_out && ctl && j + NST < mine) load_tile(j + NST);
        if (warp < 16) epilogue(j);
      } else {
        wait_mma(j);
        if (ctl && j + NST < mine) load_tile(j + NST);
        if (warp < 16) epilogue(j);
        if (j + 1 < mine) {
          __syncthreads();                               // every warp has drained the accumulators of tile j
          if (ctl) issue_mma(j + 1);
        }
      }
    }
    if (stage_out) {
      __syncthreads();
      if (ctl) {
        store_tile(mine - 1);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}


// ------------------------------------------------------------------------------------
// wide groups (CG > 96: stages 3 - 4 of pvig_s, CG = 200 / 320; arch 't' stage 3, CG = 120)
// ------------------------------------------------------------------------------------
// Four accumulators of ceil16(CG) columns no longer fit the tensor memory, nor four weight blocks the shared memory.
// A work item is (conv group q, column pass, 128-row tile): the A sub-tile of group q (128 x KP) and the weight block
// of the pass (NT x KP, NT <= 160 output channels) sit in shared memory, one accumulator of NT columns in TMEM.  Items
// are ordered (q, pass)-major, so a CTA reloads its weight block only when it moves to the next (q, pass): ~8 times per
// launch.  Same arithmetic and epilogue as the kernel above.
struct WideParams {
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  const __nv_bfloat16* w_op;     // [4][passes][NT/8][KP/8][8][8]
  const float* shift;
  long long rows;
  int C2, CG, KP, NT, passes;
};

template <int ACT>
__global__ void __launch_bounds__(THREADS, 2) grouped_fc_wide_kernel(const WideParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int CG = prm.CG, KP = prm.KP, NT = prm.NT, C2 = prm.C2;
  const uint32_t a_bytes = (uint32_t)BM * KP * 2, b_bytes = (uint32_t)NT * KP * 2;
  uint8_t* sA = smem;
  uint8_t* sB = sA + a_bytes;
  float* s_shift = reinterpret_cast<float*>(sB + b_bytes);          // [NT] of the current (q, pass)
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_shift + NT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = NT <= 128 ? 128u : 256u;

  for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t sbo = (uint32_t)(KP >> 3) * 128u;
  const int cg_chunks = CG >> 3;
  const long long tiles = (prm.rows + BM - 1) / BM;
  const long long items = tiles * 4 * prm.passes;
  uint32_t phase = 0;
  int cur_qp = -1;

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int qp = (int)(item / tiles);
    const long long tile = item - (long long)qp * tiles;
    const int q = qp / prm.passes, pass = qp - q * prm.passes;
    const long long r0 = tile * BM;
    if (qp != cur_qp) {            // new weight block + shifts (the previous item's MMAs have retired: barrier below)
      const uint4* wsrc = reinterpret_cast<const uint4*>(prm.w_op) + (size_t)qp * (b_bytes / 16);
      for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += THREADS) reinterpret_cast<uint4*>(sB)[i] = __ldg(wsrc + i);
      for (int i = threadIdx.x; i < NT; i += THREADS) {
        const int o = pass * NT + i;
        s_shift[i] = o < CG ? prm.shift[q * CG + o] : 0.f;
      }
      cur_qp = qp;
    }
    for (int i = threadIdx.x; i < BM * cg_chunks; i += THREADS) {
      const int r = i / cg_chunks, c8 = i - r * cg_chunks;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r0 + r < prm.rows) v = __ldg(reinterpret_cast<const uint4*>(prm.in + (r0 + r) * C2 + q * CG) + c8);
      *reinterpret_cast<uint4*>(sA + ((uint32_t)((r >> 3) * (KP >> 3) + c8) << 7) + ((r & 7) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      for (int ks = 0; ks < (KP >> 4); ++ks) {
        const uint64_t ad = make_desc(a_addr + ks * 256, 128, sbo), bd = make_desc(b_addr + ks * 256, 128, sbo);
        const uint32_t acc = ks != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(smem_u32(bar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      // warp w reads TMEM lanes 32*(w%4).., columns (w/4) * 16, + 64, ... (four warps per lane quarter)
      const int row = (warp & 3) * 32 + lane;
      const bool row_ok = r0 + row < prm.rows;
      __nv_bfloat16* orow = prm.out + (r0 + row) * C2 + q * CG + pass * NT;
      const int ncol = min(NT, CG - pass * NT);            // valid output channels of this pass (multiple of 8)
      for (int c0 = (warp >> 2) * 16; c0 < ncol; c0 += 64) {
        uint32_t acc[16];
        tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, acc);
        __align__(16) __nv_bfloat162 o[8];
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (c0 + j4 < ncol) {
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + j4);
            v[0] = activate<ACT>(__uint_as_float(acc[j4]) + sh.x);
            v[1] = activate<ACT>(__uint_as_float(acc[j4 + 1]) + sh.y);
            v[2] = activate<ACT>(__uint_as_float(acc[j4 + 2]) + sh.z);
            v[3] = activate<ACT>(__uint_as_float(acc[j4 + 3]) + sh.w);
          }
          o[j4 >> 1] = __floats2bfloat162_rn(v[0], v[1]);
          o[(j4 >> 1) + 1] = __floats2bfloat162_rn(v[2], v[3]);
        }
        if (row_ok) {
          *reinterpret_cast<uint4*>(orow + c0) = *reinterpret_cast<const uint4*>(o);
          if (c0 + 8 < ncol) *reinterpret_cast<uint4*>(orow + c0 + 8) = *reinterpret_cast<const uint4*>(o + 4);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------
// weight gradient: gw[q][o][i] = sum_r go[r, q*CG + o] * x[r, q*CG + i]
// ------------------------------------------------------------------------------------
// Per conv group a (CG x R) . (R x CG) product with the reduction over ALL rows: both operands are read exactly as
// they lie in memory -- row r holds the CG channels of a group contiguously, which is the UMMA "MN-major" operand
// form (channel index fastest, K = row index strided), so no transposition is needed on the way to shared memory.
// A CTA owns (group q, 128-channel slab of output channels, column pass of <= 256 input channels, row split s): it
// walks its rows in blocks of 64 (K = 64 per block, 4 MMAs of M = 128, N = NT), accumulates in TMEM and adds its
// partial (fp32) into gw with atomics at the end (split-R reduction; gw must be zero-filled by the caller).
constexpr int WG_KB = 64;
constexpr int WG_STAGES = 3;
struct WgradParams {
  const __nv_bfloat16* go;       // (rows, C2)
  const __nv_bfloat16* x;        // (rows, C2)
  float* gw;                     // (4, CG, CG) fp32, zero-filled
  long long rows;
  int C2, CG, NT, mtiles, passes, splits;
  long long rows_per_split;      // multiple of WG_KB
};

__global__ void __launch_bounds__(256, 2) grouped_fc_wgrad_kernel(const WgradParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int CG = prm.CG, NT = prm.NT, C2 = prm.C2;
  constexpr uint32_t kSbo = (WG_KB / 8) * 128;                       // bytes between groups of 8 channels
  const uint32_t a_bytes = 16 * kSbo, b_bytes = (uint32_t)(NT / 8) * kSbo;
  uint8_t* sA = smem;                                               // WG_STAGES x [16 channel groups][8 row groups][8 rows][8 channels]
  uint8_t* sB = sA + WG_STAGES * a_bytes;                           // WG_STAGES x [NT/8 channel groups][...]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + WG_STAGES * b_bytes);   // [WG_STAGES]: MMAs of a buffer retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + WG_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int it = blockIdx.x;
  const int s = it % prm.splits; it /= prm.splits;
  const int pass = it % prm.passes; it /= prm.passes;
  const int mt = it % prm.mtiles;
  const int q = it / prm.mtiles;
  const int o0 = mt * 128, i0 = pass * NT;                          // first output / input channel of this CTA (in the group)
  const long long rbeg = (long long)s * prm.rows_per_split;
  const long long rend = min(prm.rows, rbeg + prm.rows_per_split);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) mbar_init(smem_u32(bar + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NT >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const int a_groups = min(16, (CG - o0 + 7) / 8), b_groups = min(NT / 8, (CG - i0 + 7) / 8);
  // Three-stage cp.async ring over blocks of WG_KB rows: block i+2 is in flight while block i is multiplied (the
  // synchronous version paid one DRAM round trip per 64-row block: ~100 us per call whatever the layer).
  uint32_t ph[WG_STAGES] = {0, 0, 0};
  const long long nblocks = rend > rbeg ? (rend - rbeg + WG_KB - 1) / WG_KB : 0;
  const int nblk = (int)nblocks;
  auto load_block = [&](long long blk) {            // asynchronous; one commit group per call, empty past the end
    if (blk < nblocks) {
      const long long r0 = rbeg + blk * WG_KB;
      const uint32_t a = smem_u32(sA) + (uint32_t)(blk % WG_STAGES) * a_bytes;
      const uint32_t b = smem_u32(sB) + (uint32_t)(blk % WG_STAGES) * b_bytes;
      // 16-byte pieces (8 channels of one row): piece (group g, row r) -> g * kSbo + (r / 8) * 128 + (r % 8) * 16.
      // Narrow operands (<= 8 channel groups): only the groups that exist are copied (CG = 40 fills 5 of the 16 groups of the
      // M = 128 operand: the padding groups are zeroed once, below) and 8 consecutive lanes take the 8 rows of one core matrix (128 contiguous
      // bytes of shared memory: lanes walking the groups of ONE row all hit the same banks, 1 KB apart).
      if (a_groups <= 8) {
        for (int i = threadIdx.x; i < a_groups * WG_KB; i += 256) {
          const int r8 = i & 7, t = i >> 3, g = t % a_groups, r = (t / a_groups) * 8 + r8;
          const bool ok = r0 + r < rend;
          cp_async16(a + g * kSbo + ((r >> 3) << 7) + (r8 << 4),
                     ok ? reinterpret_cast<const uint4*>(prm.go + (r0 + r) * C2 + q * CG + o0) + g
                        : reinterpret_cast<const uint4*>(prm.go), ok ? 16u : 0u);
        }
      } else {                                      // wide slabs: 256 contiguous bytes per row (measured faster there)
        for (int i = threadIdx.x; i < 16 * WG_KB; i += 256) {
          const int g = i % 16, r = i / 16;
          const bool ok = g < a_groups && r0 + r < rend;
          cp_async16(a + g * kSbo + ((r >> 3) << 7) + ((r & 7) << 4),
                     ok ? reinterpret_cast<const uint4*>(prm.go + (r0 + r) * C2 + q * CG + o0) + g
                        : reinterpret_cast<const uint4*>(prm.go), ok ? 16u : 0u);
        }
      }
      if (b_groups <= 8) {
        for (int i = threadIdx.x; i < b_groups * WG_KB; i += 256) {
          const int r8 = i & 7, t = i >> 3, g = t % b_groups, r = (t / b_groups) * 8 + r8;
          const bool ok = r0 + r < rend;
          cp_async16(b + g * kSbo + ((r >> 3) << 7) + (r8 << 4),
                     ok ? reinterpret_cast<const uint4*>(prm.x + (r0 + r) * C2 + q * CG + i0) + g
                        : reinterpret_cast<const uint4*>(prm.x), ok ? 16u : 0u);
        }
      } else {
        for (int i = threadIdx.x; i < (NT / 8) * WG_KB; i += 256) {
          const int g = i % (NT / 8), r = i / (NT / 8);
          const bool ok = g < b_groups && r0 + r < rend;
          cp_async16(b + g * kSbo + ((r >> 3) << 7) + ((r & 7) << 4),
                     ok ? reinterpret_cast<const uint4*>(prm.x + (r0 + r) * C2 + q * CG + i0) + g
                        : reinterpret_cast<const uint4*>(prm.x), ok ? 16u : 0u);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // channel groups past the end of the conv group stay zero in every stage
  {
    const uint32_t a_pad = (uint32_t)(16 - a_groups) * kSbo / 16, b_pad = (uint32_t)(NT / 8 - b_groups) * kSbo / 16;   // uint4 per stage
    for (uint32_t i = threadIdx.x; i < WG_STAGES * a_pad; i += 256)
      reinterpret_cast<uint4*>(sA + (i / a_pad) * a_bytes + (uint32_t)a_groups * kSbo)[i % a_pad] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < WG_STAGES * b_pad; i += 256)
      reinterpret_cast<uint4*>(sB + (i / b_pad) * b_bytes + (uint32_t)b_groups * kSbo)[i % b_pad] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }
  load_block(0);
  load_block(1);
  for (long long blk = 0; blk < nblocks; ++blk) {
    const int buf = (int)(blk % WG_STAGES);
    asm volatile("cp.async.wait_group 1;" ::: "memory");           // block blk has landed (blk + 1 may still be in flight)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(sA) + buf * a_bytes, b_addr = smem_u32(sB) + buf * b_bytes;
      for (int ks = 0; ks < WG_KB / 16; ++ks) {                      // K = 16 rows = two row groups = 256 bytes
        const uint64_t ad = make_desc(a_addr + ks * 256, 128, kSbo), bd = make_desc(b_addr + ks * 256, 128, kSbo);
        const uint32_t acc = (blk | ks) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + buf)) : "memory");
    }
    // block blk + 2 goes into the buffer block blk - 1 used: its MMAs (issued an iteration ago) must have retired
    if (blk >= 1) {
      const int pb = (int)((blk - 1) % WG_STAGES);
      mbar_wait(smem_u32(bar + pb), ph[pb]);
      ph[pb] ^= 1;
    }
    load_block(blk + 2);
  }
  // drain: the commit of the last block covers every earlier MMA
  if (nblk > 0) {
    const int lb = (int)((nblocks - 1) % WG_STAGES);
    mbar_wait(smem_u32(bar + lb), ph[lb]);
    ph[lb] ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4 && nblk > 0) {
    const int o = o0 + warp * 32 + lane;                             // TMEM lane == output channel
    float* dst = prm.gw + ((size_t)q * CG + o) * CG + i0;
    const int ncol = min(NT, CG - i0);
    for (int c0 = 0; c0 < ncol; c0 += 16) {
      uint32_t acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      if (o < CG) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < ncol) atomicAdd(dst + c0 + j, __uint_as_float(acc[j]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

}  // namespace fc
}  // namespace gkg

using namespace gkg;

// narrow groups: the four accumulators share the tensor memory and the four weight blocks the shared memory
static bool fc_narrow_ok(int CG) {
  const int KP = (CG + 15) / 16 * 16;
  const size_t smem = 4 * (size_t)fc::BM * KP * 2 + 4 * (size_t)KP * KP * 2 + (size_t)4 * CG * 4 + 64;
  return 4 * KP <= 512 && smem <= 200 * 1024 && CG <= 96;
}
struct WidePlan { int KP, NT, passes; size_t smem; bool ok; };
static WidePlan fc_wide_plan(int CG) {
  WidePlan p{};
  p.KP = (CG + 15) / 16 * 16;
  p.passes = (p.KP + 159) / 160;
  p.NT = ((p.KP + p.passes - 1) / p.passes + 15) / 16 * 16;
  p.smem = (size_t)fc::BM * p.KP * 2 + (size_t)p.NT * p.KP * 2 + (size_t)p.NT * 4 + 64;
  p.ok = CG >= 8 && p.NT <= 256 && p.smem <= 220 * 1024;
  return p;
}

extern "C" int gkg_grouped_fc_supported(int C2) {
  if (C2 <= 0 || C2 % 32) return 0;                     // 4 conv groups of a multiple of 8 channels
  const int CG = C2 / 4;
  return (fc_narrow_ok(CG) || fc_wide_plan(CG).ok) ? 1 : 0;
}

// layout of the weight operand for this width: 0 = [4][NP/8][KP/8][8][8] (narrow), else the column-pass width NT of the
// wide layout [4][passes][NT/8][KP/8][8][8]
extern "C" int gkg_grouped_fc_pass_width(int C2) {
  if (!gkg_grouped_fc_supported(C2)) return -1;
  const int CG = C2 / 4;
  return fc_narrow_ok(CG) ? 0 : fc_wide_plan(CG).NT;
}

// 3-D tensor map over a (rows, C2) bf16 activation: (8 channels, rows, C2 / 8), box (8, BM, C2 / 8); rows past the end read as
// zeros.  false when the driver entry point is missing or rejects the shape (the caller then takes the cp.async kernel).
static bool fc_input_map(CUtensorMap* map, const void* in, long long rows, int C2) {
  static const PFN_cuTensorMapEncodeTiled_v12000 encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }();
  if (encode == nullptr || C2 / 8 > 256) return false;
  const cuuint64_t gdim[3] = {8, (cuuint64_t)rows, (cuuint64_t)(C2 / 8)};
  const cuuint64_t gstr[2] = {(cuuint64_t)C2 * 2, 16};
  const cuuint32_t box[3] = {8, (cuuint32_t)fc::BM, (cuuint32_t)(C2 / 8)};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(in), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int gkg_grouped_fc_fwd(const void* in, const void* w_op, const float* shift, void* out, long long rows,
                                  int C2, int act, gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(rows >= 0 && gkg_grouped_fc_supported(C2), "grouped_fc_fwd: unsupported shape rows=%lld 2C=%d", rows, C2);
  GKG_CHECK_ARG(act >= 0 && act <= 2, "grouped_fc_fwd: bad activation %d", act);
  if (rows == 0) return GKG_OK;
  GKG_CHECK_ARG(in && w_op && shift && out, "grouped_fc_fwd: null pointer");
  GKG_CHECK_ARG(((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)w_op % 16) == 0,
                "grouped_fc_fwd: pointers must be 16-byte aligned");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (rows + fc::BM - 1) / fc::BM;
  const int CG = C2 / 4;
  if (!fc_narrow_ok(CG)) {
    const WidePlan wp = fc_wide_plan(CG);
    fc::WideParams prm{};
    prm.in = static_cast<const __nv_bfloat16*>(in);
    prm.out = static_cast<__nv_bfloat16*>(out);
    prm.w_op = static_cast<const __nv_bfloat16*>(w_op);
    prm.shift = shift; prm.rows = rows;
    prm.C2 = C2; prm.CG = CG; prm.KP = wp.KP; prm.NT = wp.NT; prm.passes = wp.passes;
    void (*kern)(const fc::WideParams) = act == 0 ? fc::grouped_fc_wide_kernel<0>
                                       : act == 1 ? fc::grouped_fc_wide_kernel<1> : fc::grouped_fc_wide_kernel<2>;
    static std::atomic<uint64_t> configured[3];
    cudaError_t e = cudaSuccess;
    configure_once_per_device(configured[act], [&] {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    if (e != cudaSuccess) { set_error("grouped_fc_fwd: smem attribute: %s", cudaGetErrorString(e)); return GKG_ECUDA; }
    const long long items = tiles * 4 * wp.passes;
    // two co-resident CTAs per SM when their shared memory allows (CG = 120, 200): one copies rows in while the
    // other is in its MMA / epilogue (the kernel itself is synchronous)
    const long long ctas = (long long)sms * (wp.smem <= 110 * 1024 ? 2 : 1);
    const int grid = (int)(items < ctas ? items : ctas);
    kern<<<grid, fc::THREADS, wp.smem, stream>>>(prm);
    GKG_CHECK_LAUNCH("grouped_fc_wide_kernel");
    return GKG_OK;
  }
  fc::Params prm{};
  prm.in = static_cast<const __nv_bfloat16*>(in);
  prm.out = static_cast<__nv_bfloat16*>(out);
  prm.w_op = static_cast<const __nv_bfloat16*>(w_op);
  prm.shift = shift; prm.rows = rows;
  prm.C2 = C2; prm.CG = C2 / 4; prm.KP = (prm.CG + 15) / 16 * 16; prm.NP = prm.KP; prm.act = act;
  const size_t smem = 4 * (size_t)fc::BM * prm.KP * 2 + 4 * (size_t)prm.NP * prm.KP * 2 + (size_t)C2 * 4 + 64;
  // TMA copy-in (grouped_fc_tma_kernel) when the driver encodes the 3-D map and at least two stages fit
  if (tiles >= 2LL * sms && rows < (1LL << 31)) {
    const size_t tile_bytes = (size_t)fc::BM * C2 * 2;
    const size_t fixed = 2048 + 4 * (size_t)prm.NP * prm.KP * 2 + (size_t)C2 * 4 + 8 * 6 + 16;
    int nst = 4;
    while (nst >= 2 && nst * tile_bytes + fixed > 220 * 1024) --nst;
    CUtensorMap tmap;
    if (nst >= 2 && fc_input_map(&tmap, in, rows, C2)) {
      prm.stages = nst;
      prm.stage_out = (8 * prm.NP <= 512 && nst >= 3) ? 1 : 0;
      void (*tk)(const fc::Params, const CUtensorMap) = nullptr;
#define GKG_FC_PICK(AA)                                                                                        \
      tk = prm.CG == 40 ? fc::grouped_fc_tma_kernel<40, AA> : prm.CG == 80 ? fc::grouped_fc_tma_kernel<80, AA> : fc::grouped_fc_tma_kernel<0, AA>
      if (act == 0) { GKG_FC_PICK(0); } else if (act == 1) { GKG_FC_PICK(1); } else { GKG_FC_PICK(2); }
#undef GKG_FC_PICK
      const int tslot = act * 3 + (prm.CG == 40 ? 0 : prm.CG == 80 ? 1 : 2);
      static std::atomic<uint64_t> tconfigured[9];
      cudaError_t e = cudaSuccess;
      configure_once_per_device(tconfigured[tslot], [&] {
        e = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      });
      if (e != cudaSuccess) { set_error("grouped_fc_fwd: smem attribute: %s", cudaGetErrorString(e)); return GKG_ECUDA; }
      const int tgrid = (int)(tiles < sms ? tiles : sms);
      tk<<<tgrid, fc::THREADS + 32, nst * tile_bytes + fixed, stream>>>(prm, tmap);
      GKG_CHECK_LAUNCH("grouped_fc_tma_kernel");
      return GKG_OK;
    }
  }
  const size_t smem_pipe = smem + 4 * (size_t)fc::BM * prm.KP * 2;          // second A buffer
  if (smem_pipe <= 220 * 1024 && tiles >= 2LL * sms) {
    // software-pipelined kernel: one persistent CTA per SM (worth it once every SM has a few tiles)
    void (*pk)(const fc::Params) = nullptr;
#define GKG_FC_PICK(AA)                                                                                        \
    pk = prm.CG == 40 ? fc::grouped_fc_pipe_kernel<40, AA> : prm.CG == 80 ? fc::grouped_fc_pipe_kernel<80, AA> : fc::grouped_fc_pipe_kernel<0, AA>
    if (act == 0) { GKG_FC_PICK(0); } else if (act == 1) { GKG_FC_PICK(1); } else { GKG_FC_PICK(2); }
#undef GKG_FC_PICK
    const int pslot = act * 3 + (prm.CG == 40 ? 0 : prm.CG == 80 ? 1 : 2);
    static std::atomic<uint64_t> pconfigured[9];
    cudaError_t e = cudaSuccess;
    configure_once_per_device(pconfigured[pslot], [&] {
      e = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    if (e != cudaSuccess) { set_error("grouped_fc_fwd: smem attribute %zu: %s", smem_pipe, cudaGetErrorString(e)); return GKG_ECUDA; }
    const int pgrid = (int)(tiles < sms ? tiles : sms);
    pk<<<pgrid, fc::THREADS, smem_pipe, stream>>>(prm);
    GKG_CHECK_LAUNCH("grouped_fc_pipe_kernel");
    return GKG_OK;
  }
  void (*kern)(const fc::Params) = nullptr;
  int slot = act * 3;
#define GKG_FC_PICK(AA)                                                                                        \
  kern = prm.CG == 40 ? fc::grouped_fc_kernel<40, AA> : prm.CG == 80 ? fc::grouped_fc_kernel<80, AA> : fc::grouped_fc_kernel<0, AA>
  if (act == 0) { GKG_FC_PICK(0); } else if (act == 1) { GKG_FC_PICK(1); } else { GKG_FC_PICK(2); }
#undef GKG_FC_PICK
  slot += prm.CG == 40 ? 0 : prm.CG == 80 ? 1 : 2;
  {
    static std::atomic<uint64_t> configured[9];
    cudaError_t e = cudaSuccess;
    configure_once_per_device(configured[slot], [&] {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    if (e != cudaSuccess) { set_error("grouped_fc_fwd: smem attribute %zu: %s", smem, cudaGetErrorString(e)); return GKG_ECUDA; }
  }
  const int ctas_per_sm = (4 * prm.NP <= 256 && smem <= 110 * 1024) ? 2 : 1;
  const int grid = (int)(tiles < (long long)sms * ctas_per_sm ? tiles : (long long)sms * ctas_per_sm);
  kern<<<grid, fc::THREADS, smem, stream>>>(prm);
  GKG_CHECK_LAUNCH("grouped_fc_kernel");
  return GKG_OK;
}

// Weight operand packing in one launch: (4, CG, CG) fp32 conv weight [q][o][i] (optionally scaled per output channel,
// optionally transposed per group for the data gradient) -> bf16 core-matrix order of the forward kernels.
__global__ void grouped_fc_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                       __nv_bfloat16* __restrict__ out, int CG, int KP, int NT, int passes, int transpose) {
  const int nrows = NT > 0 ? passes * NT : KP;                     // padded output rows per group
  const long long total = 4LL * nrows * KP;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    // destination index -> (q, n, k): [q][(pass)][n/8][k/8][n%8][k%8]
    const int k8 = (int)(e & 7), n8 = (int)((e >> 3) & 7);
    long long t = e >> 6;
    const int kc = (int)(t % (KP >> 3)); t /= (KP >> 3);
    const int rows8 = (NT > 0 ? NT : KP) >> 3;
    const int nc = (int)(t % rows8); t /= rows8;
    int n = nc * 8 + n8;
    if (NT > 0) { const int pass = (int)(t % passes); t /= passes; n += pass * NT; }
    const int q = (int)t, k = kc * 8 + k8;
    float v = 0.f;
    if (n < CG && k < CG) {
      v = transpose ? w[((size_t)q * CG + k) * CG + n] : w[((size_t)q * CG + n) * CG + k];
      if (scale != nullptr) v *= scale[q * CG + n];
    }
    out[e] = __float2bfloat16_rn(v);
  }
}

extern "C" int gkg_grouped_fc_pack_weights(const float* weight, const float* scale, void* w_op, int C2, int transpose,
                                           gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(gkg_grouped_fc_supported(C2), "grouped_fc_pack_weights: unsupported width 2C=%d", C2);
  GKG_CHECK_ARG(weight && w_op, "grouped_fc_pack_weights: null pointer");
  const int CG = C2 / 4, KP = (CG + 15) / 16 * 16;
  int NT = 0, passes = 1;
  if (!fc_narrow_ok(CG)) { const WidePlan wp = fc_wide_plan(CG); NT = wp.NT; passes = wp.passes; }
  const long long total = 4LL * (NT > 0 ? passes * NT : KP) * KP;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  grouped_fc_pack_kernel<<<blocks, 256, 0, stream>>>(weight, scale, static_cast<__nv_bfloat16*>(w_op), CG, KP, NT, passes,
                                                     transpose);
  GKG_CHECK_LAUNCH("grouped_fc_pack_kernel");
  return GKG_OK;
}

extern "C" int gkg_grouped_fc_wgrad(const void* grad_out, const void* in, float* grad_w, long long rows, int C2,
                                    gkg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GKG_CHECK_ARG(rows >= 0 && C2 > 0 && C2 % 32 == 0 && C2 / 4 <= 1024, "grouped_fc_wgrad: unsupported shape rows=%lld 2C=%d", rows, C2);
  if (rows == 0) return GKG_OK;
  GKG_CHECK_ARG(grad_out && in && grad_w, "grouped_fc_wgrad: null pointer");
  GKG_CHECK_ARG(((uintptr_t)grad_out % 16) == 0 && ((uintptr_t)in % 16) == 0, "grouped_fc_wgrad: pointers must be 16-byte aligned");
  const int CG = C2 / 4;
  fc::WgradParams prm{};
  prm.go = static_cast<const __nv_bfloat16*>(grad_out);
  prm.x = static_cast<const __nv_bfloat16*>(in);
  prm.gw = grad_w; prm.rows = rows; prm.C2 = C2; prm.CG = CG;
  const int NP = (CG + 15) / 16 * 16;
  prm.passes = (NP + 255) / 256;
  prm.NT = ((NP + prm.passes - 1) / prm.passes + 15) / 16 * 16;
  prm.mtiles = (CG + 127) / 128;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int base = 4 * prm.mtiles * prm.passes;
  long long blocks_of_rows = (rows + fc::WG_KB - 1) / fc::WG_KB;
  // row splits: every split ends in CG x CG fp32 atomics, so the wide groups (160 k addresses, tens of atomics each)
  // take one CTA per SM; the narrow ones two (their epilogue is small, their copy-in latency needs the overlap)
  int splits = ((CG > 96 ? 1 : 2) * sms + base - 1) / base;
  if (splits > blocks_of_rows) splits = (int)blocks_of_rows;
  if (splits < 1) splits = 1;
  prm.rows_per_split = (blocks_of_rows + splits - 1) / splits * fc::WG_KB;
  splits = (int)((rows + prm.rows_per_split - 1) / prm.rows_per_split);
  prm.splits = splits;
  const size_t smem = fc::WG_STAGES * (size_t)(16 + prm.NT / 8) * (fc::WG_KB / 8) * 128 + 64;
  static std::atomic<uint64_t> configured{0};
  cudaError_t e = cudaSuccess;
  configure_once_per_device(configured, [&] {
    e = cudaFuncSetAttribute(fc::grouped_fc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (e != cudaSuccess) { set_error("grouped_fc_wgrad: smem attribute: %s", cudaGetErrorString(e)); return GKG_ECUDA; }
  fc::grouped_fc_wgrad_kernel<<<base * splits, 256, smem, stream>>>(prm);
  GKG_CHECK_LAUNCH("grouped_fc_wgrad_kernel");
  return GKG_OK;
}

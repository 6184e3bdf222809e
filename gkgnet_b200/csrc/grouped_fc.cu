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
#include "common.cuh"

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
template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return 0.5f * v * (1.f + fast_erf(v * 0.70710678118654752f));    // nn.GELU() (erf form)
  return v;
}

struct Params {
  const __nv_bfloat16* in;       // (rows, C2) contiguous
  __nv_bfloat16* out;            // (rows, C2) contiguous
  const __nv_bfloat16* w_op;     // 4 groups x [NP/8][KP/8][8][8] core matrices (K-major)
  const float* shift;            // (C2); the per-channel scale is folded into w_op by the caller
  long long rows;
  int C2, CG, KP, NP, act;
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

}  // namespace fc
}  // namespace gkg

using namespace gkg;

extern "C" int gkg_grouped_fc_supported(int C2) {
  if (C2 <= 0 || C2 % 32) return 0;                     // 4 conv groups of a multiple of 8 channels
  const int CG = C2 / 4, NP = (CG + 15) / 16 * 16;
  return 4 * NP <= 512 ? 1 : 0;                         // the four accumulators must fit the TMEM columns
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
  fc::Params prm{};
  prm.in = static_cast<const __nv_bfloat16*>(in);
  prm.out = static_cast<__nv_bfloat16*>(out);
  prm.w_op = static_cast<const __nv_bfloat16*>(w_op);
  prm.shift = shift; prm.rows = rows;
  prm.C2 = C2; prm.CG = C2 / 4; prm.KP = (prm.CG + 15) / 16 * 16; prm.NP = prm.KP; prm.act = act;
  const size_t smem = 4 * (size_t)fc::BM * prm.KP * 2 + 4 * (size_t)prm.NP * prm.KP * 2 + (size_t)C2 * 4 + 64;
  void (*kern)(const fc::Params) = nullptr;
#define GKG_FC_PICK(AA)                                                                                        \
  kern = prm.CG == 40 ? fc::grouped_fc_kernel<40, AA> : prm.CG == 80 ? fc::grouped_fc_kernel<80, AA> : fc::grouped_fc_kernel<0, AA>
  if (act == 0) { GKG_FC_PICK(0); } else if (act == 1) { GKG_FC_PICK(1); } else { GKG_FC_PICK(2); }
#undef GKG_FC_PICK
  {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("grouped_fc_fwd: smem attribute %zu: %s", smem, cudaGetErrorString(e)); return GKG_ECUDA; }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (rows + fc::BM - 1) / fc::BM;
  const int ctas_per_sm = (4 * prm.NP <= 256 && smem <= 110 * 1024) ? 2 : 1;
  const int grid = (int)(tiles < (long long)sms * ctas_per_sm ? tiles : (long long)sms * ctas_per_sm);
  kern<<<grid, fc::THREADS, smem, stream>>>(prm);
  GKG_CHECK_LAUNCH("grouped_fc_kernel");
  return GKG_OK;
}

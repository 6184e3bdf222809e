// Device side of the tcgen05 kNN kernel (see knn_tc.cu for the design notes).  Included by
// knn_tc.cu (host logic) and knn_tc_inst.cu (one translation unit per bias mode, so the
// template instantiations compile in parallel).
#pragma once
#include "knn_tc.cuh"

namespace gkg {
namespace tc {


constexpr int BM = 128;            // query rows per tile  (UMMA M)
constexpr int BN = 144;            // keys per tile        (UMMA N): 4 chunks of 36 = lcm of the
                                   // key-grid widths 9/18/36 of the separable position bias
constexpr int CH = 36;             // columns per epilogue chunk (tcgen05.ld x32 + x4)
constexpr int NTHREADS = 192;      // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int NACC = 3;            // TMEM accumulators
constexpr int ACC_STRIDE = 160;    // TMEM columns between accumulators (3 x 160 <= 512)
constexpr int SEP_B_FLOATS = 2048; // staged rows of the separable bias table B (512 per epilogue warp)
constexpr int CAND_SLACK = 24;     // per-row candidate buffer = T survivors + CAND_SLACK new entries
constexpr int MAX_T = 38;
constexpr float kScale = 256.f;    // operand scale S
constexpr float kPadKey = -60000.f;  // B extra column of padded keys -> dist ~ +468
constexpr float kDelta = 4e-6f;    // bound on |approx - exact| of the fp16x3 GEMM (dist units)
constexpr int MAX_A_BUF = 2;
constexpr int MAX_STAGES = 8;

struct Plan {
  int KP, KC, NKB, NA, NS, QT, KT;
  uint32_t a_tile_bytes, b_block_bytes;
  size_t smem_bytes;
  size_t a_op_bytes, b_op_bytes;   // per launch operand buffers
  bool ok;
};

constexpr size_t kSmemBudget = 227 * 1024;
__host__ __device__ constexpr size_t cand_bytes(int T) { return (size_t)BM * (T + CAND_SLACK) * 8; }
constexpr size_t kBarBytes = 1024 + SEP_B_FLOATS * 4;

inline Plan make_plan(int P, int N, int M, int D, int T = MAX_T) {
  const size_t kCandBytes = cand_bytes(T);
  Plan pl{};
  pl.KP = (3 * D + 2 + 15) / 16 * 16;
  pl.QT = (N + BM - 1) / BM;
  pl.KT = (M + BN - 1) / BN;
  pl.a_tile_bytes = (uint32_t)BM * pl.KP * 2;
  pl.ok = false;
  for (int na = MAX_A_BUF; na >= 1 && !pl.ok; --na) {
    const size_t fixed = kCandBytes + kBarBytes + (size_t)na * pl.a_tile_bytes;
    if (fixed >= kSmemBudget) continue;
    const size_t room = kSmemBudget - fixed;
    for (int kc = pl.KP; kc >= 16; kc -= 16) {
      if (pl.KP % kc) continue;
      const size_t blk = (size_t)BN * kc * 2;
      int ns = (int)(room / blk);
      if (ns > MAX_STAGES) ns = MAX_STAGES;
      const int want = (na == 1) ? 2 : 3;
      if (ns >= want || (ns >= 2 && kc == 16)) {
        pl.NA = na; pl.KC = kc; pl.NKB = pl.KP / kc; pl.NS = ns;
        pl.b_block_bytes = (uint32_t)blk;
        pl.ok = true;
        break;
      }
    }
  }
  if (!pl.ok) return pl;
  pl.smem_bytes = kCandBytes + kBarBytes + (size_t)pl.NA * pl.a_tile_bytes + (size_t)pl.NS * pl.b_block_bytes;
  pl.a_op_bytes = (size_t)P * pl.QT * pl.a_tile_bytes;
  pl.b_op_bytes = (size_t)P * pl.KT * (size_t)BN * pl.KP * 2;
  return pl;
}

// ------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread (no issue slots) until the
// phase completes or ~hint ns pass.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must abort the kernel, never hang the GPU.
// BACKOFF: single-lane producer / MMA warps sleep between polls so that their spinning does not
// take issue slots from the epilogue warp sharing the scheduler.
template <bool BACKOFF>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!(BACKOFF ? mbar_try_wait_hint(bar, parity, 20000u) : mbar_try_wait(bar, parity))) {
    if (BACKOFF) __nanosleep(200);
    if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// One epilogue chunk = 36 accumulator columns of this thread's row: x32 + x4 loads.
__device__ __forceinline__ void tmem_ld36(uint32_t taddr, uint32_t (&r)[36]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35])
               : "r"(taddr + 32)
               : "memory");
}
// The loaded registers are threaded through the wait so the compiler cannot hoist their uses.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[36]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]),
        "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35])
      :
      : "memory");
}

// UMMA shared-memory descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes stored
// contiguously; LBO = byte distance between core matrices adjacent in K, SBO = between 8-row
// groups (cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at 17, M>>4 at 24.
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// Sorted list of the T smallest VALUES seen so far (registers).  Branch-free insertion:
// new[s] = min(max(x, old[s-1]), old[s]) -- two FMNMX per slot, independent across slots, a
// no-op when x >= v[T-1].  The (value, id) pairs themselves live in the row's shared-memory
// buffer: [0, ns) survivors (value <= tau at the last compaction), [ns, cnt) new candidates.
template <int T>
struct ValList {
  float v[T];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < T; ++s) v[s] = INFINITY;
  }
  __device__ __forceinline__ void insert(float x) {
#pragma unroll
    for (int s = T - 1; s >= 1; --s) v[s] = fminf(fmaxf(x, v[s - 1]), v[s]);
    v[0] = fminf(x, v[0]);
  }
};

// Fold the new candidates of every lane into its value list, then keep only the entries that can
// still be among the T smallest.  Loop trip counts are warp-uniform; bodies are predicated.
template <int T>
__device__ __forceinline__ void compact_candidates(ValList<T>& top, float& tau, float2*& wp, int& ns,
                                                   bool& overflow, float2* cbuf) {
  const int cnt = (int)(wp - cbuf) / BM;
  const int mx_new = __reduce_max_sync(0xffffffffu, cnt - ns);
  for (int i = 0; i < mx_new; ++i) {
    const int e = ns + i;
    const float x = (e < cnt) ? cbuf[e * BM].x : INFINITY;
    top.insert(x);
  }
  tau = top.v[T - 1];
  const int mx_all = __reduce_max_sync(0xffffffffu, cnt);
  int w = 0;
  for (int e = 0; e < mx_all; ++e) {
    if (e < cnt) {
      const float2 c = cbuf[e * BM];
      if (c.x <= tau) {
        cbuf[w * BM] = c;
        ++w;
      }
    }
  }
  if (w > T + 4) {          // many exact ties at the threshold: keep room, certify later by fix-up
    w = T + 4;
    overflow = true;
  }
  ns = w;
  wp = cbuf + w * BM;
}

struct TcParams {
  const __half* a_op;
  const __half* b_op;
  const float* xhat; const float* xsq; const float* yhat; const float* ysq;
  const float* relpos;             // dense (N, M) bias, or null
  const float* sep_a;              // separable bias: A (grid_w, KW), B (N / grid_w, M / KW)
  const float* sep_b;
  int grid_w, sep_mh;
  int32_t* idx_out;
  int* fix_count; int* fix_rows; unsigned int* stats;   // stats: [0] ambiguous rows, [1] max err bits
  float* dbg_dist;
  int P, N, M, D, k, dilation, kd;
  int KP, KC, NKB, NA, NS, QT, KT;
  uint32_t a_tile_bytes, b_block_bytes;
  int force_rerank;
};

__device__ __forceinline__ float exact_dist(const float* __restrict__ xr, const float* __restrict__ yr, int D,
                                            float xs, float ys, const float* relrow, int m) {
  float acc = 0.f;
  for (int d = 0; d < D; ++d) acc = fmaf(xr[d], yr[d], acc);
  float v = (xs + (-2.f * acc)) + ys;
  if (relrow != nullptr) v += relrow[m];
  return v;
}

// BIAS: 0 = none, 1 = dense relative_pos read per element, KW (9 / 18 / 36) = separable
// bias  relpos[n, m] = A[n % grid_w][m % KW] + B[n / grid_w][m / KW]  with the A row in registers
// and the needed B rows staged in shared memory (the analytic table of the reference has this
// form: pos_embed.py + the flattened bicubic resize, see gkgnet_b200/pos_embed.py).
template <int T, int BIAS>
__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_kernel(const TcParams prm) {
  constexpr bool HAS_REL = BIAS != 0;
  constexpr bool DENSE = BIAS == 1;
  constexpr int KW = BIAS > 1 ? BIAS : 36;
  extern __shared__ __align__(1024) uint8_t smem[];
  // carve-up: [A x NA][B ring x NS][candidates][barriers + tmem ptr][staged B rows]
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)prm.NA * prm.a_tile_bytes;
  float2* cand = reinterpret_cast<float2*>(sB + (size_t)prm.NS * prm.b_block_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(cand) + cand_bytes(T));
  uint64_t* a_full = bars;                    // [MAX_A_BUF]
  uint64_t* a_empty = a_full + MAX_A_BUF;     // [MAX_A_BUF]
  uint64_t* b_full = a_empty + MAX_A_BUF;     // [MAX_STAGES]
  uint64_t* b_empty = b_full + MAX_STAGES;    // [MAX_STAGES]
  uint64_t* t_full = b_empty + MAX_STAGES;    // [NACC]
  uint64_t* t_empty = t_full + NACC;          // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + NACC);
  float* sepB_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // [SEP_B_FLOATS]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_A_BUF; ++i) { mbar_init(smem_u32(a_full + i), 1); mbar_init(smem_u32(a_empty + i), 1); }
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(smem_u32(b_full + i), 1); mbar_init(smem_u32(b_empty + i), 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(smem_u32(t_full + i), 1); mbar_init(smem_u32(t_empty + i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_items = prm.P * prm.QT;

  if (warp == 0) {
    // ================================ TMA producer ===================================
    if (lane == 0) {
      int ab = 0, aph = 0, bs = 0, bph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int p = item / prm.QT, qt = item - p * prm.QT;
        mbar_wait<true>(smem_u32(a_empty + ab), aph ^ 1);
        mbar_expect_tx(smem_u32(a_full + ab), prm.a_tile_bytes);
        tma_bulk_g2s(smem_u32(sA + (size_t)ab * prm.a_tile_bytes),
                     reinterpret_cast<const uint8_t*>(prm.a_op) + ((size_t)p * prm.QT + qt) * prm.a_tile_bytes,
                     prm.a_tile_bytes, smem_u32(a_full + ab));
        if (++ab == prm.NA) { ab = 0; aph ^= 1; }
        const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(prm.b_op) +
                              (size_t)p * prm.KT * prm.NKB * prm.b_block_bytes;
        const int nblk = prm.KT * prm.NKB;
        for (int blk = 0; blk < nblk; ++blk) {
          mbar_wait<true>(smem_u32(b_empty + bs), bph ^ 1);
          mbar_expect_tx(smem_u32(b_full + bs), prm.b_block_bytes);
          tma_bulk_g2s(smem_u32(sB + (size_t)bs * prm.b_block_bytes), bsrc + (size_t)blk * prm.b_block_bytes,
                       prm.b_block_bytes, smem_u32(b_full + bs));
          if (++bs == prm.NS) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =====================================
    if (lane == 0) {
      int ab = 0, aph = 0, bs = 0, bph = 0, tb = 0, tph = 0;
      const uint32_t lbo = 128, sbo = (uint32_t)(prm.KC >> 3) * 128;
      const int ksteps = prm.KC >> 4;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        mbar_wait<true>(smem_u32(a_full + ab), aph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + (size_t)ab * prm.a_tile_bytes);
        for (int kt = 0; kt < prm.KT; ++kt) {
          mbar_wait<true>(smem_u32(t_empty + tb), tph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)tb * ACC_STRIDE;
          for (int kb = 0; kb < prm.NKB; ++kb) {
            mbar_wait<true>(smem_u32(b_full + bs), bph);
            tc_fence_after();
            const uint32_t a_addr = a_base + (uint32_t)kb * (BM * prm.KC * 2);
            const uint32_t b_addr = smem_u32(sB + (size_t)bs * prm.b_block_bytes);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t ad = make_smem_desc(a_addr + ks * 256, lbo, sbo);
              const uint64_t bd = make_smem_desc(b_addr + ks * 256, lbo, sbo);
              umma_f16(d_tmem, ad, bd, kIdesc, (kb | ks) != 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(b_empty + bs));       // frees the B block when the MMAs retire
            if (++bs == prm.NS) { bs = 0; bph ^= 1; }
          }
          umma_commit(smem_u32(t_full + tb));           // accumulator ready for the epilogue
          if (++tb == NACC) { tb = 0; tph ^= 1; }
        }
        umma_commit(smem_u32(a_empty + ab));            // A tile may be overwritten
        if (++ab == prm.NA) { ab = 0; aph ^= 1; }
      }
    }
  } else {
    // ================================ epilogue / selection ===========================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int row_t = q * 32 + lane;
    float2* cbuf = cand + row_t;                 // entry e at cbuf[e * BM]
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float c_scale = -2.f / (kScale * kScale);
    int tb = 0, tph = 0;
    ValList<T> top;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int p = item / prm.QT, qt = item - p * prm.QT;
      const int n = qt * BM + row_t;
      const bool row_ok = n < prm.N;
      const int n_c = row_ok ? n : prm.N - 1;
      const float* relrow = HAS_REL ? prm.relpos + (size_t)n_c * prm.M : nullptr;
      top.init();
      float tau = INFINITY;
      float2* wp = cbuf;                           // next free candidate slot of this row
      int ns = 0;                                  // survivors at the front of the buffer
      bool overflow = false;

      // ---- separable bias: A row -> registers, B rows of this warp's 32 rows -> shared memory
      float areg[KW];
      const float* brow = sepB_s;
      if (BIAS > 1) {
        float* mine = sepB_s + q * (SEP_B_FLOATS / 4);
        const int first = min(prm.N - 1, qt * BM + q * 32);
        const int last = min(prm.N - 1, qt * BM + q * 32 + 31);
        const int h0 = first / prm.grid_w;
        const int nh = last / prm.grid_w - h0 + 1;
        __syncwarp();
        for (int i = lane; i < nh * prm.sep_mh; i += 32) mine[i] = __ldg(prm.sep_b + (size_t)h0 * prm.sep_mh + i);
        __syncwarp();
        const float* arow = prm.sep_a + (size_t)(n_c % prm.grid_w) * KW;
#pragma unroll
        for (int j = 0; j < KW; ++j) areg[j] = __ldg(arow + j);
        brow = mine + (n_c / prm.grid_w - h0) * prm.sep_mh;
      }

      float bias[DENSE ? CH : 1];
      auto load_bias = [&](int m0) {
        if (DENSE) {
          if ((prm.M & 3) == 0) {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (m0 + j4 * 4 < prm.M) b4 = __ldg(reinterpret_cast<const float4*>(relrow + m0 + j4 * 4));
              bias[j4 * 4 + 0] = b4.x; bias[j4 * 4 + 1] = b4.y; bias[j4 * 4 + 2] = b4.z; bias[j4 * 4 + 3] = b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) bias[DENSE ? j : 0] = (m0 + j < prm.M) ? __ldg(relrow + m0 + j) : 0.f;
          }
        }
      };

      for (int kt = 0; kt < prm.KT; ++kt) {
        mbar_wait<false>(smem_u32(t_full + tb), tph);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / CH; ++c) {
          uint32_t r[CH];
          tmem_ld36(lane_addr + (uint32_t)(tb * ACC_STRIDE + c * CH), r);
          const int m0 = kt * BN + c * CH;
          load_bias(m0);
          tmem_ld_wait(r);
          if (prm.dbg_dist != nullptr && row_ok) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              if (m0 + j < prm.M) {
                const float acc = __uint_as_float(r[j]);
                float b = 0.f;
                if (DENSE) b = bias[DENSE ? j : 0];
                if (BIAS > 1) b = areg[j % KW] + brow[min(m0 / KW + j / KW, prm.sep_mh - 1)];
                prm.dbg_dist[((size_t)p * prm.N + n) * prm.M + m0 + j] = fmaf(acc, c_scale, b);
              }
            }
          }
#pragma unroll
          for (int g = 0; g < CH / KW; ++g) {
            // per key group: fold the B term into the threshold, add it back on the rare pass
            float bg = 0.f;
            if (BIAS > 1) bg = brow[min(m0 / KW + g, prm.sep_mh - 1)];
            const float taug = tau - bg;
#pragma unroll
            for (int jj = 0; jj < KW; ++jj) {
              const int j = g * KW + jj;
              const float acc = __uint_as_float(r[j]);
              float v;
              if (DENSE) v = fmaf(acc, c_scale, bias[DENSE ? j : 0]);
              else if (BIAS > 1) v = fmaf(acc, c_scale, areg[jj]);
              else v = acc * c_scale;
              if (v < taug) {
                *wp = make_float2(v + bg, __int_as_float(m0 + j));
                wp += BM;
              }
              // the buffer must always have room for the next 12 candidates
              if ((j % 12) == 11 && __any_sync(0xffffffffu, wp > cbuf + (T + CAND_SLACK - 12) * BM)) {
                compact_candidates<T>(top, tau, wp, ns, overflow, cbuf);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(t_empty + tb));
        if (++tb == NACC) { tb = 0; tph ^= 1; }
      }
      compact_candidates<T>(top, tau, wp, ns, overflow, cbuf);

      // ---------------- finalise the row -------------------------------------------
      // survivors -> registers, sorted by (value, id)
      float cv[T];
      int cid[T];
#pragma unroll
      for (int s = 0; s < T; ++s) {
        const float2 c = cbuf[s * BM];
        const bool have = s < ns;
        cv[s] = have ? c.x : INFINITY;
        cid[s] = have ? __float_as_int(c.y) : 0x7fffffff;
      }
      auto sort_pairs = [&]() {
#pragma unroll
        for (int pass = 0; pass < T; ++pass) {
#pragma unroll
          for (int s = pass & 1; s + 1 < T; s += 2) {
            const bool sw = (cv[s + 1] < cv[s]) || (cv[s + 1] == cv[s] && cid[s + 1] < cid[s]);
            const float tv = sw ? cv[s] : cv[s + 1];
            const int ti = sw ? cid[s] : cid[s + 1];
            cv[s] = sw ? cv[s + 1] : cv[s];
            cid[s] = sw ? cid[s + 1] : cid[s];
            cv[s + 1] = tv;
            cid[s + 1] = ti;
          }
        }
      };
      sort_pairs();

      const int kd = prm.kd;
      bool amb = prm.force_rerank != 0 || overflow || ns > T;
#pragma unroll
      for (int s = 0; s + 1 < T; ++s)
        if (s < kd && (cv[s + 1] - cv[s]) < 2.f * kDelta) amb = true;
      if (row_ok && amb) {
        // exact fp32 re-rank of the T candidates (identical arithmetic to knn_exact.cu)
        const float* xr = prm.xhat + ((size_t)p * prm.N + n) * prm.D;
        const float xs = prm.xsq[(size_t)p * prm.N + n];
        const float* yb = prm.yhat + (size_t)p * prm.M * prm.D;
        const float* ysb = prm.ysq + (size_t)p * prm.M;
        const float a_last = tau;                  // every non-candidate has approx >= tau
        float maxerr = 0.f;
#pragma unroll
        for (int s = 0; s < T; ++s) {
          const int m = cid[s];
          if (m < prm.M) {
            const float e = exact_dist(xr, yb + (size_t)m * prm.D, prm.D, xs, ysb[m], relrow, m);
            maxerr = fmaxf(maxerr, fabsf((e - xs) - cv[s]));
            cv[s] = e;
          } else {
            cv[s] = INFINITY;
          }
        }
        sort_pairs();
        atomicAdd(prm.stats + 0, 1u);
        atomicMax(prm.stats + 1, __float_as_uint(maxerr));
        float e_kd = -INFINITY;               // sorted ascending: kd-th value == max of the first kd
#pragma unroll
        for (int s = 0; s < T; ++s)
          if (s < kd) e_kd = fmaxf(e_kd, cv[s]);
        const bool unsure = overflow || ns > T || (a_last - kDelta <= (e_kd - xs) + kDelta);
        if (unsure && prm.M > T) {
          const int slot = atomicAdd(prm.fix_count, 1);
          prm.fix_rows[slot] = p * prm.N + n;
        }
      }
      // stage the ids in this thread's (now idle) candidate slots so that the dilated pick is a
      // shared-memory index, not a dynamic register index
#pragma unroll
      for (int s = 0; s < T; ++s) cbuf[s * BM].y = __int_as_float(cid[s]);
      if (row_ok) {
        int32_t* out = prm.idx_out + ((size_t)p * prm.N + n) * prm.k;
        for (int j = 0; j < prm.k; ++j) out[j] = __float_as_int(cbuf[j * prm.dilation * BM].y);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}


// launcher for one bias mode; instantiated in knn_tc_inst.cu
template <int BIAS>
int launch_select(const TcParams& prm, const Plan& pl, int T, cudaStream_t stream);

}  // namespace tc
}  // namespace gkg

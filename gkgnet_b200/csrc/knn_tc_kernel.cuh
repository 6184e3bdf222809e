// Device side of the tcgen05 kNN kernel (see knn_tc.cu for the design notes).  Included by
// knn_tc.cu (host logic) and knn_tc_inst.cu (one translation unit per bias mode, so the
// template instantiations compile in parallel).
#pragma once
#include "knn_tc.cuh"

namespace gkg {
namespace tc {

constexpr int BM = 128;            // query rows per row set (UMMA M)
constexpr int CH = 36;             // accumulator columns per epilogue chunk (tcgen05.ld x32 + x4) = lcm of
                                   // the key-grid widths 9/18/36 of the separable position bias
constexpr int MAX_T = 38;
constexpr float kScale = 256.f;    // operand scale S
constexpr float kNegHalfS2 = -0.5f * 256.f * 256.f;   // beta = kNegHalfS2 * relative_pos (exact: power of two)
constexpr float kScoreToDist = 1.f / kNegHalfS2;      // dist - |xh|^2 = score * kScoreToDist
constexpr float kPadKey = -60000.f;  // B extra column of padded keys -> score ~ -1.5e7 (dist ~ +468)
constexpr float kScoreFloor = -8e6f; // initial threshold: above every padded key, below every real one (dist < 244)
// Bound on |approx - exact| of the fp16x3 GEMM in distance units: every K=16 accumulation step rounds
// the fp32 accumulator (|acc| <= 1.5 S^2 -> 3.6e-7 per step), plus the 2^-21 relative split error.
inline float tc_delta(int KP) { return 3.6e-7f * (float)(KP / 16) + 1.2e-6f; }
inline float tc_delta_steps(int ksteps) { return 3.6e-7f * (float)ksteps + 1.2e-6f; }
// |single-plane - fp16x3| <= S^2 * 2^-10 * sum|xh_i yh_i| <= 64 (Cauchy-Schwarz, unit rows); + fp32 accumulation
constexpr float kSweepSlack = 66.f;
constexpr int MAX_STAGES = 8;
constexpr size_t kSmemBudget = 227 * 1024;

// Kernel geometry.  One work item = RS row sets of 128 queries of one problem against all its keys;
// every B block that streams through shared memory feeds RS MMAs, so RS = 2 halves the operand
// traffic per query and doubles the epilogue warps (latency hiding) at the price of a log per row.
//   BN   real keys per key tile (multiple of CH)      BNP  UMMA N (BN rounded up to 16; pad rows never read)
//   NACC accumulator slots per row set
template <int RS_, int BN_, int BNP_, int NACC_, int ACC_STRIDE_, int SEPW_>
struct Geom {
  static constexpr int RS = RS_, BN = BN_, BNP = BNP_, NACC = NACC_, ACC_STRIDE = ACC_STRIDE_;
  static constexpr int NCH = BN / CH;             // chunks per accumulator
  static constexpr int NEPI = 4 * RS;             // epilogue warps
  // MMA issuer warps: one per row set; with a single row set, NACC of them take the key tiles (= accumulator
  // slots) in turn -- the issue loop of ONE warp (~130 instructions per operand block: barrier wait, descriptors,
  // commits) was what bounded the wide-group shapes, not the tensor pipe
  static constexpr int NISS = RS == 1 ? NACC_ : RS;
  static constexpr int NTHREADS = 32 * (1 + NISS + NEPI); // warp 0 TMA, warps 1..NISS MMA issue, then the epilogue warps
  static constexpr int ROWS = BM * RS;
  static constexpr uint32_t LOG_STRIDE = ROWS * 16;   // bytes between consecutive log slots of a row
  static constexpr int SEPW = SEPW_;              // staged floats of the separable bias table B per epilogue warp
  static constexpr size_t kBarBytes = 1024 + (size_t)NEPI * SEPW * 4;
  // kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at 17, M>>4 at 24.
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BNP >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(BN % CH == 0 && BNP % 16 == 0 && BNP >= BN && RS * NACC * ACC_STRIDE <= 512 && ACC_STRIDE >= BNP, "geometry");
  // Triplet log of a row: log_valid(T) entries can be kept (the second sweep logs ~T + 1 on average, see
  // the kernel); behind them log_slack(T) slots absorb the triplets logged between two capacity checks
  // (6 per half chunk).  T + 8 for every list length (it was T + T / 2 for the long lists: the 12 KB that frees at
  // T = 29 buy the wide-group plans a fourth operand stage, -4 % at D = 200, k*d = 27; a row that logs more compacts).
  __host__ __device__ static constexpr int log_valid(int T) { return T + 8; }
  __host__ __device__ static constexpr int log_slack(int T) { return 6; }
  __host__ __device__ static constexpr size_t cand_bytes(int T) {
    return (size_t)ROWS * (log_valid(T) + log_slack(T)) * 16;
  }
};
using GeomA = Geom<1, 144, 144, 3, 160, 512>;   // 128 rows / item: any shape
using GeomB = Geom<2, 72, 80, 3, 80, 384>;      // 256 rows / item: small operands (D <= 80), short lists

// K layout of the operands (see tc_prepare_*): [x_hi (D) | e_hi e_lo | 0.. -> PA | second (D) | third (D) | 0.. -> KP]
// so that the first PA columns alone give the single-plane product S^2 (x_hi.y_hi - |yh|^2/2) of sweep A.
__host__ __device__ constexpr int k_prefix(int D) { return (D + 2 + 15) / 16 * 16; }
__host__ __device__ constexpr int k_padded(int D) { return k_prefix(D) + (2 * D + 15) / 16 * 16; }

struct Plan {
  int geom;                        // 0 = GeomA, 1 = GeomB
  int KP, PA, KC, NKB, NKBA, NA, NS, QT, QI, QTP, KT;
  int split, KS1, KSL;             // split-plane operands (see below): valid K steps of the first / the lo segment
  uint32_t a_tile_bytes, a_res_bytes, b_block_bytes, stage_bytes;
  size_t smem_bytes;
  size_t a_op_bytes, b_op_bytes;   // per launch operand buffers
  bool ok;
};

// Three operand-staging modes:
//   resident  (NA = 1 | 2 buffers per row set): the whole 128-row A tiles (all K) sit in shared memory
//             for the item, B blocks (BNP keys x KC) stream through the ring;
//   split     (wide groups, D % 8 == 0, GeomA): the fp16x3 operand rows [hi | hi | lo] / [hi | lo | hi] repeat
//             their hi plane, so only TWO planes per row are stored, [hi (+ extra pair) -> PA | lo -> KP - PA] for
//             queries and keys alike (each segment zero-padded to a multiple of KC).  The hi segment of the A tile
//             stays resident; a ring stage carries one B block with the matching A.lo block (first segment in
//             sweep B), or else TWO consecutive B blocks (they are adjacent in global memory: one bulk copy).  The issuer multiplies  A.hi x B.first (+ A.lo x B.first in sweep B)  for
//             a block of the first segment and  A.hi x B.lo  for a block of the lo segment: per key tile 233 KB
//             through the ring instead of 444 KB at D = 200;
//   streaming (NA = 0):     A and B blocks of one K slice travel together through the ring (the A
//             tiles would not leave room for the triplet log) -- A is re-read once per key tile.
template <class G>
inline Plan make_plan_g(int P, int N, int M, int D, int T, bool strict) {
  Plan pl{};
  pl.KP = k_padded(D);
  pl.PA = k_prefix(D);
  pl.QT = (N + BM - 1) / BM;
  pl.QI = (pl.QT + G::RS - 1) / G::RS;
  pl.QTP = pl.QI * G::RS;
  pl.KT = (M + G::BN - 1) / G::BN;
  pl.a_tile_bytes = (uint32_t)BM * pl.KP * 2;
  pl.ok = false;
  const size_t cand = G::cand_bytes(T);
  if (pl.KP <= 256) {
    // preference order: a well fed pipeline first (>= 3 B stages of a decent size), A double-buffered if it fits
    for (int pass = 0; pass < (strict ? 1 : 2) && !pl.ok; ++pass) {
      for (int na = 2; na >= 1 && !pl.ok; --na) {
        const size_t fixed = cand + G::kBarBytes + (size_t)na * G::RS * pl.a_tile_bytes;
        if (fixed >= kSmemBudget) continue;
        const size_t room = kSmemBudget - fixed;
        for (int kc = pl.KP; kc >= 16; kc -= 16) {
          if (pl.KP % kc) continue;
          const size_t blk = (size_t)G::BNP * kc * 2;
          int ns = (int)(room / blk);
          if (ns > MAX_STAGES) ns = MAX_STAGES;
          const int want = (na == 1) ? 2 : 3;
          const bool good = ns >= 3 && (size_t)ns * blk >= 40 * 1024;
          const bool usable = ns >= want || (ns >= 2 && kc == 16);
          if (pass == 0 ? good : usable) {
            pl.NA = na; pl.KC = kc; pl.NS = ns;
            pl.b_block_bytes = (uint32_t)blk;
            pl.ok = true;
            break;
          }
        }
      }
    }
  }
#ifndef GKG_NO_SPLIT
  if (!pl.ok && !strict && G::RS == 1 && D % 8 == 0) {
    // pass 0: the widest blocks of >= 32 columns with >= 4 stages and >= 64 KB in flight; pass 1: >= 3 stages and
    // >= 48 KB; pass 2: whatever keeps the most bytes in flight.  (Measured at D = 200: 5 x 17 KB beats 3 x 26 KB
    // and 8 x 9 KB for lists of 20; 3 x 17 KB beats 8 x 9 KB for lists of 29 -- fewer barrier round trips per tile.)
    size_t best = 0;
    for (int pass = 0; pass < 3 && !pl.ok; ++pass) {
      for (int kc = 64; kc >= 16; kc -= 16) {
#ifdef GKG_SPLIT_KC                                   // experiments only: force the block width of the split plan
        if (kc != GKG_SPLIT_KC) continue;
#endif
        const int pas = (D + 2 + kc - 1) / kc * kc, pls = (D + kc - 1) / kc * kc;
        const size_t a_res = (size_t)BM * pas * 2;
        const size_t fixed = cand + G::kBarBytes + a_res;
        if (fixed >= kSmemBudget) continue;
        const size_t stage = (size_t)(2 * G::BNP > G::BNP + BM ? 2 * G::BNP : G::BNP + BM) * kc * 2;   // B + A.lo | 2 B
        int ns = (int)((kSmemBudget - fixed) / stage);
        if (ns > MAX_STAGES) ns = MAX_STAGES;
        if (ns < 3) continue;
        const size_t flight = (size_t)ns * stage;
        const bool take = pass == 0 ? (kc >= 32 && ns >= 4 && flight >= 64 * 1024)
                        : pass == 1 ? (kc >= 32 && flight >= 48 * 1024)
                                    : flight > best;
        if (!take) continue;
        best = flight;
        pl.split = 1; pl.NA = 1; pl.KC = kc; pl.NS = ns;
        pl.PA = pas; pl.KP = pas + pls;
        pl.b_block_bytes = (uint32_t)((size_t)G::BNP * kc * 2);
        pl.a_tile_bytes = (uint32_t)((size_t)BM * (pas + pls) * 2);
        if (pass < 2) { pl.ok = true; break; }
      }
      if (pass == 2 && best > 0) pl.ok = true;
    }
    if (pl.ok) {
      pl.KS1 = (D + 2 + 15) / 16;
      pl.KSL = (D + 15) / 16;
    }
  }
#endif
  if (!pl.ok && !strict && cand + G::kBarBytes < kSmemBudget) {
    const size_t room = kSmemBudget - cand - G::kBarBytes;
    for (int kc = 64; kc >= 16 && !pl.ok; kc -= 16) {
      if (pl.KP % kc) continue;
      const size_t blk = (size_t)(G::ROWS + G::BNP) * kc * 2;
      int ns = (int)(room / blk);
      if (ns > MAX_STAGES) ns = MAX_STAGES;
      if (ns >= 4 || (ns >= 2 && kc == 16)) {
        pl.NA = 0; pl.KC = kc; pl.NS = ns;
        pl.b_block_bytes = (uint32_t)((size_t)G::BNP * kc * 2);
        pl.ok = true;
      }
    }
  }
  if (!pl.ok) return pl;
  pl.NKB = pl.KP / pl.KC;
  pl.NKBA = (pl.PA + pl.KC - 1) / pl.KC;     // split: the blocks of the first segment
  const size_t stage = pl.split ? (size_t)(2 * G::BNP > G::BNP + BM ? 2 * G::BNP : G::BNP + BM) * pl.KC * 2
                                : pl.b_block_bytes + (pl.NA == 0 ? (size_t)G::ROWS * pl.KC * 2 : 0);
  pl.stage_bytes = (uint32_t)stage;
  pl.a_res_bytes = pl.split ? (uint32_t)((size_t)BM * pl.PA * 2) : pl.a_tile_bytes;   // resident part of an A tile
  pl.smem_bytes = cand + G::kBarBytes + (size_t)pl.NA * G::RS * pl.a_res_bytes + (size_t)pl.NS * stage;
  pl.a_op_bytes = (size_t)P * pl.QTP * pl.a_tile_bytes;
  pl.b_op_bytes = (size_t)P * pl.KT * (size_t)G::BNP * pl.KP * 2;
  return pl;
}

// GeomB when its resident plan is comfortable (small operands, list of <= 20, enough rows to fill the
// row sets), GeomA otherwise.
inline Plan make_plan(int P, int N, int M, int D, int T = MAX_T) {
  if (T <= 20 && N > BM) {
    Plan pb = make_plan_g<GeomB>(P, N, M, D, T, true);
    if (pb.ok) { pb.geom = 1; return pb; }
  }
  Plan pa = make_plan_g<GeomA>(P, N, M, D, T, false);
  pa.geom = 0;
  return pa;
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
// BACKOFF: the single-lane producer / MMA warps nap briefly between polls so that their spinning
// does not take issue slots from the epilogue warp sharing the scheduler.  (A long suspend-time
// hint on try_wait was measured to wake up microseconds late; short naps are the better trade.)
template <bool BACKOFF>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (BACKOFF) __nanosleep(32);
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}
// Single-lane variant for the lean producer / issuer loops below: no timer reads on the fast path, a spin bound
// instead of a wall-clock bound (2^26 naps of >= 20 ns: more than a second).
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(20);
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One lane of the (converged) warp: the issuer of TMA / tcgen05 instructions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

// ------------------------------------------------------------------------------------
// selection machinery (epilogue warps; one thread == one query row)
// ------------------------------------------------------------------------------------
// Everything is ranked in the SCALED SCORE domain  s = acc + beta,  acc = S^2 (xh.yh - |yh|^2/2)
// straight out of TMEM and beta = -(S^2/2) * relative_pos, so that
//     dist - |xh|^2 = s * (-2 / S^2)          (both factors are powers of two: exact)
// and "nearest" == "largest score".  The keys of an item are swept TWICE:
//   sweep A  (tensor cores: the first PA operand columns only = single-plane fp16 product, 3/8 of the
//            MMA work at D = 40) keeps, per row, the TA = T - 1 largest maxima of 18-key groups in a
//            sorted register list (branch-free insertion).  They are TA distinct keys, so the TA-th
//            largest score of the row is >= tauA - kSweepSlack, where kSweepSlack bounds the
//            difference between the single-plane and the fp16x3 product;
//   sweep B  (all KP columns = fp16x3 product) logs every TRIPLET of keys (three scores + the id of
//            the first, one predicated 16-byte shared-memory store) whose maximum beats that fixed
//            threshold: ~T + 1 triplets per row, no list maintenance, no log compaction.
// The final selection then works on the logged triplets only.
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // FMNMX3

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 ld_shared_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}

// The T largest triplet maxima seen so far, sorted descending, in registers.  Branch-free
// insertion: new[s] = max(old[s], min(x, old[s-1])) -- two FMNMX per slot, a no-op when x <= v[T-1].
template <int T>
struct TopList {
  float v[T];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < T; ++s) v[s] = kScoreFloor;
  }
  __device__ __forceinline__ void insert(float x) {
#pragma unroll
    for (int s = T - 1; s >= 1; --s) v[s] = fmaxf(fminf(x, v[s - 1]), v[s]);
    v[0] = fmaxf(x, v[0]);
  }
};

// Per-row triplet log in shared memory: entry e of the row handled by thread t lives at
// log_base(t) + e * LOG_STRIDE (16 bytes per thread, consecutive lanes adjacent -> conflict-free
// 128-bit accesses).
// T-th largest KEY score among the `cnt` log entries of this thread's row (kScoreFloor when the row has logged
// fewer than T keys); end of the item only: it converts the log to true scores on the way.  Triplet maxima first; the other two keys of a triplet are inserted only when
// they can matter for some row of the warp (neighbouring keys are often similar).  Warp-uniform trip counts.
template <uint32_t LOG_STRIDE, int T, int BIAS>
__device__ __forceinline__ float log_tth_key(uint32_t log_base, int cnt, int mx_cnt, const float* brow) {
  TopList<T> top;
  top.init();
  // first pass: the B term of the key group is added back IN PLACE (from here on the log holds true scores)
  for (int e0 = 0; e0 < mx_cnt; e0 += 2) {
    float4 c[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) c[u] = ld_shared_v4(log_base + (uint32_t)(e0 + u) * LOG_STRIDE);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool valid = e0 + u < cnt;
      if (BIAS > 1) {
        const float bg = brow[valid ? (__float_as_uint(c[u].w) >> 16) : 0u];
        c[u].x += bg; c[u].y += bg; c[u].z += bg;
        if (valid) st_shared_v4(log_base + (uint32_t)(e0 + u) * LOG_STRIDE, c[u].x, c[u].y, c[u].z, c[u].w);
      }
      top.insert(valid ? fmax3(c[u].x, c[u].y, c[u].z) : kScoreFloor);
    }
  }
  for (int e = 0; e < mx_cnt; ++e) {
    float x1 = kScoreFloor, x2 = kScoreFloor;
    if (e < cnt) {
      const float4 c = ld_shared_v4(log_base + (uint32_t)e * LOG_STRIDE);
      const float mx = fmax3(c.x, c.y, c.z);
      const bool is0 = c.x == mx, is1 = !is0 && c.y == mx;      // the one copy of the maximum already listed
      x1 = is0 ? c.y : c.x;
      x2 = (is0 || is1) ? c.z : c.y;
    }
    if (__any_sync(0xffffffffu, fmaxf(x1, x2) > top.v[T - 1])) {
      top.insert(x1);
      top.insert(x2);
    }
  }
  return top.v[T - 1];
}

// Rare path of sweep B: a row is past its LC valid log entries because its sweep-A threshold was loose.
// Raise the threshold to the T-th largest TRIPLET maximum logged so far (T distinct keys reach it; a key
// that ties with it must stay a candidate: a hair of slack) and drop the entries below, which leaves at
// most T (+ ties) of them.  Small on purpose: it is inlined at every capacity check of the chunk loop.
template <uint32_t LOG_STRIDE, int T, int BIAS>
__device__ __forceinline__ void log_compact(uint32_t log_base, uint32_t& lp, float& thr_base, const float* brow) {
  const int cnt_now = (int)((lp - log_base) / LOG_STRIDE);
  const int mx_now = __reduce_max_sync(0xffffffffu, cnt_now);
  TopList<T> top;
  top.init();
#pragma unroll 1
  for (int e = 0; e < mx_now; ++e) {
    const float4 c = ld_shared_v4(log_base + (uint32_t)e * LOG_STRIDE);
    const float bg = BIAS > 1 ? brow[e < cnt_now ? (__float_as_uint(c.w) >> 16) : 0u] : 0.f;
    top.insert(e < cnt_now ? fmax3(c.x, c.y, c.z) + bg : kScoreFloor);
  }
  thr_base = fmaxf(thr_base, top.v[T - 1] - 1.f);
  uint32_t wp = log_base;
#pragma unroll 1
  for (int e = 0; e < mx_now; ++e) {
    const float4 c = ld_shared_v4(log_base + (uint32_t)e * LOG_STRIDE);
    const float bg = BIAS > 1 ? brow[e < cnt_now ? (__float_as_uint(c.w) >> 16) : 0u] : 0.f;
    if (e < cnt_now && fmax3(c.x, c.y, c.z) + bg > thr_base) {
      st_shared_v4(wp, c.x, c.y, c.z, c.w);
      wp += LOG_STRIDE;
    }
  }
  lp = wp;
}

// One chunk of a query row in flight: CH accumulator columns and the bias terms that go with them.
template <bool DENSE, int NG>
struct Chunk {
  uint32_t r[CH];
  float bias[DENSE ? CH : 1];
  float bg[NG];
};

struct TcParams {
  const __half* a_op;
  const __half* b_op;
  const float* yhat; const float* ysq;
  const float* relpos;             // dense (N, M) bias, or null
  const float* sep_a;              // separable bias: A (grid_w, KW), B (N / grid_w, M / KW)
  const float* sep_b;
  int grid_w, sep_mh, sep_mhp;
  int32_t* idx_out;
  int* fix_count; int* fix_rows; unsigned int* stats;   // stats: [0] ambiguous rows, [1] max err bits
  int* rr_count; int* rr_list; int rr_cap;              // rows whose candidates go to the exact re-rank kernel
  float2* cand; int* cand_count; float* cand_thr; int cand_slots;   // per item: [slot][row] (score, id), [row] count / threshold
  float* dbg_dist;
  int P, N, M, D, k, dilation, kd;
  int KP, PA, KC, NKB, NKBA, NA, NS, QT, QI, QTP, KT;
  int split, KS1, KSL;
  uint32_t a_tile_bytes, a_res_bytes, b_block_bytes, stage_bytes;
  int force_rerank;
  float delta;                     // bound on |approx - exact| of the fp16x3 GEMM (dist units), see tc_delta()
};

// exact distance in the reference's association order (knn_exact.cu); xr: the normalised query row (shared memory)
__device__ __forceinline__ float exact_dist(const float* __restrict__ xr, const float* __restrict__ yr, int D,
                                            float xs, float ys, const float* relrow, int m) {
  float acc = 0.f;
  if ((D & 3) == 0) {      // same fma chain, 128-bit loads (rows start on 16-byte boundaries when D % 4 == 0)
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    const float4* y4 = reinterpret_cast<const float4*>(yr);
#pragma unroll 4
    for (int d = 0; d < (D >> 2); ++d) {
      const float4 a = x4[d], b = __ldg(y4 + d);
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
  } else {
    for (int d = 0; d < D; ++d) acc = fmaf(xr[d], __ldg(yr + d), acc);
  }
  float v = (xs + (-2.f * acc)) + ys;
  if (relrow != nullptr) v += __ldg(relrow + m);
  return v;
}

// BIAS: 0 = none, 1 = dense relative_pos read per element, KW (9 / 18 / 36) = separable
// bias  relpos[n, m] = A[n % grid_w][m % KW] + B[n / grid_w][m / KW]  with the A row in registers
// and the needed B rows staged in shared memory (the analytic table of the reference has this
// form: pos_embed.py + the flattened bicubic resize, see gkgnet_b200/pos_embed.py).
// DBG: the instantiation that can dump the raw distance matrix (tests); kept out of the production kernel,
// where the dump code alone was a third of the sweep-B loop body (instruction cache).
template <class G, int T, int BIAS, int GA, bool DBG>
__global__ void __launch_bounds__(G::NTHREADS, 1) knn_tc_kernel(const TcParams prm) {
  constexpr bool HAS_REL = BIAS != 0;
  constexpr bool DENSE = BIAS == 1;
  constexpr int KW = BIAS > 1 ? BIAS : CH;      // columns that share one B term (whole chunk if none)
  constexpr int TA = T - 1;                     // sweep-A list: TA >= k*d + 1 distinct keys
  constexpr int LC = G::log_valid(T);           // log entries a row may keep
  constexpr int NCH = G::NCH;                   // chunks per accumulator
  constexpr int NG = CH / KW;                   // key groups per chunk
  constexpr int RS = G::RS, NACC = G::NACC;
  constexpr uint32_t LOG_STRIDE = G::LOG_STRIDE;
  static_assert(KW == 9 || KW == 18 || KW == 36, "bias period");
  extern __shared__ __align__(1024) uint8_t smem[];
  // carve-up: [A tiles x NA x RS][ring x NS: B block (+ A slices when streaming)][triplet log][barriers + tmem ptr][staged B rows]
  const uint32_t a_blk_bytes = (uint32_t)(BM * prm.KC * 2);                    // one K slice of one A tile
  const uint32_t stage_bytes = prm.stage_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)prm.NA * RS * prm.a_res_bytes;
  uint8_t* cand = sB + (size_t)prm.NS * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(cand + G::cand_bytes(T));
  uint64_t* a_full = bars;                    // [2 * RS]
  uint64_t* a_empty = a_full + 4;             // [2 * RS]
  uint64_t* b_full = a_empty + 4;             // [MAX_STAGES]
  uint64_t* b_empty = b_full + MAX_STAGES;    // [MAX_STAGES]
  uint64_t* t_full = b_empty + MAX_STAGES;    // [RS * NACC]
  uint64_t* t_empty = t_full + 8;             // [RS * NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 8);
  float* sepB_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // [NEPI][SEPW]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(a_full + i), 1); mbar_init(smem_u32(a_empty + i), RS == 1 ? G::NISS : 1); }
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(smem_u32(b_full + i), 1); mbar_init(smem_u32(b_empty + i), G::NISS); }
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(t_full + i), 1); mbar_init(smem_u32(t_empty + i), 4); }
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

  const int total_items = prm.P * prm.QI;

  if (warp == 0) {
    // ================================ TMA producer ===================================
    // The whole warp walks the loops (converged control flow); one elected lane issues the copies.
    // Per item: sweep A streams the first NKBA K blocks of every key tile, sweep B all NKB of them.
    if (RS == 1 && prm.NA > 0) {
      // Lean form for one row set with a resident A tile (every wide-group shape): loop invariants in registers,
      // addresses by increments.  The generic loop below spends ~115 instructions per operand block, which (with
      // the same weight in the issuer) bounded the D = 200 layers at ~800 cycles per 17 KB block.  The warp stays
      // converged (a divergent single-lane loop makes the compiler wrap every UBLKCP / UTCHMMA in a 15-instruction
      // uniformisation loop); one elected lane issues.
      {
        const uint32_t KT = prm.KT, NKB = prm.NKB, NKBA = prm.NKBA, NS = prm.NS, NA = prm.NA;
        const uint32_t bblk = prm.b_block_bytes, a_res = prm.a_res_bytes;
        const bool split = prm.split != 0;
        const uint32_t nkbl = NKB - NKBA;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        const uint32_t afull_u = smem_u32(a_full), aempty_u = smem_u32(a_empty);
        const uint32_t bfull_u = smem_u32(b_full), bempty_u = smem_u32(b_empty);
        uint32_t ab = 0, aph = 0, bs = 0, bph = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
          const int p = item / prm.QI, qi = item - p * prm.QI;
          const uint8_t* asrc = reinterpret_cast<const uint8_t*>(prm.a_op) +
                                ((size_t)p * prm.QTP + (size_t)qi) * prm.a_tile_bytes;
          mbar_wait_lean(aempty_u + ab * 8, aph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(afull_u + ab * 8, a_res);
            tma_bulk_g2s(sA_u + ab * a_res, asrc, a_res, afull_u + ab * 8);
          }
          __syncwarp();
          if (++ab == NA) { ab = 0; aph ^= 1; }
          const uint8_t* alo = asrc + (size_t)NKBA * a_blk_bytes;       // split: the lo segment of the A tile
          const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(prm.b_op) + (size_t)p * KT * NKB * bblk;
          for (uint32_t sweep = 0; sweep < 2; ++sweep) {
            const uint32_t nkb = sweep == 0 ? NKBA : NKB;
            const uint32_t nlo = (split && sweep == 1) ? nkbl : 0u;     // blocks that travel with an A.lo block
            const uint32_t single_end = (split && sweep == 1) ? NKBA : 0u;   // blocks [0, single_end) travel alone
            const uint8_t* src = bsrc;
            for (uint32_t kt = 0; kt < KT; ++kt, src += (size_t)NKB * bblk) {
              const uint8_t* sb = src;
              for (uint32_t kb = 0; kb < nkb;) {
                // blocks of this stage: one (with its A.lo block, if it has one), or two B blocks of one segment
                const uint32_t nb = (!split || kb < single_end) ? 1u : min(2u, nkb - kb);
                mbar_wait_lean(bempty_u + bs * 8, bph ^ 1);
                const uint32_t bar = bfull_u + bs * 8, dst = sB_u + bs * stage_bytes;
                const bool with_lo = kb < nlo;
                if (elect_one()) {
                  mbar_expect_tx(bar, nb * bblk + (with_lo ? a_blk_bytes : 0u));
                  tma_bulk_g2s(dst, sb, nb * bblk, bar);
                  if (with_lo) tma_bulk_g2s(dst + bblk, alo + (size_t)kb * a_blk_bytes, a_blk_bytes, bar);
                }
                __syncwarp();
                if (++bs == NS) { bs = 0; bph ^= 1; }
                kb += nb;
                sb += (size_t)nb * bblk;
              }
            }
          }
        }
      }
    } else {
    int ab = 0, aph = 0, bs = 0, bph = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int p = item / prm.QI, qi = item - p * prm.QI;
      const uint8_t* asrc = reinterpret_cast<const uint8_t*>(prm.a_op) +
                            ((size_t)p * prm.QTP + (size_t)qi * RS) * prm.a_tile_bytes;   // RS consecutive tiles
      if (prm.NA > 0) {
#pragma unroll
        for (int r = 0; r < RS; ++r) {
          const int slot = ab * RS + r;
          mbar_wait<true>(smem_u32(a_empty + slot), aph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(smem_u32(a_full + slot), prm.a_res_bytes);
            tma_bulk_g2s(smem_u32(sA + (size_t)slot * prm.a_res_bytes), asrc + (size_t)r * prm.a_tile_bytes,
                         prm.a_res_bytes, smem_u32(a_full + slot));
          }
          __syncwarp();
        }
        if (++ab == prm.NA) { ab = 0; aph ^= 1; }
      }
      const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(prm.b_op) +
                            (size_t)p * prm.KT * prm.NKB * prm.b_block_bytes;
      for (int sweep = 0; sweep < 2; ++sweep) {
        const int nkb = sweep == 0 ? prm.NKBA : prm.NKB;
        for (int kt = 0; kt < prm.KT; ++kt) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait<true>(smem_u32(b_empty + bs), bph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(smem_u32(b_full + bs), stage_bytes);
              uint8_t* dst = sB + (size_t)bs * stage_bytes;
              tma_bulk_g2s(smem_u32(dst), bsrc + (size_t)(kt * prm.NKB + kb) * prm.b_block_bytes, prm.b_block_bytes,
                           smem_u32(b_full + bs));
              if (prm.NA == 0) {
#pragma unroll
                for (int r = 0; r < RS; ++r)
                  tma_bulk_g2s(smem_u32(dst + prm.b_block_bytes + r * a_blk_bytes),
                               asrc + (size_t)r * prm.a_tile_bytes + (size_t)kb * a_blk_bytes, a_blk_bytes,
                               smem_u32(b_full + bs));
              }
            }
            __syncwarp();
            if (++bs == prm.NS) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
    }
  } else if (warp <= G::NISS) {
    // ================================ MMA issuers ====================================
    // RS > 1: one warp per row set, every key tile (two independent issue streams hide each other's barrier
    // latencies).  RS == 1: NACC warps, warp i issues the key tiles whose accumulator slot is i and skips the
    // operand blocks of the others.  Converged warp, one elected lane issues tcgen05.mma / tcgen05.commit (the
    // commits must come from the thread that issued the MMAs they track; elect.sync picks the same lane every time).
    constexpr bool RR = RS == 1;                 // round-robin over the key tiles
    const int iss = warp - 1;
    const int r = RR ? 0 : iss;                  // row set of this issuer
    int ab = 0, aph = 0, bs = 0, bph = 0, tb = 0, tph = 0;
    // descriptor halves: hi = SBO | version, lo = start address | LBO (both in 16-byte units)
    const uint64_t desc_hi = (uint64_t)(((uint32_t)(prm.KC >> 3) * 128u >> 4) | (1u << 14)) << 32;
    const uint32_t lbo_field = (128u >> 4) << 16;
    const int kpb = prm.KC >> 4;                 // K steps per operand block
    if (RR && prm.NA > 0) {
      // Lean form (see the producer): converged warp, invariants in registers, descriptors by addition.
      {
        const int KT = prm.KT, NKB = prm.NKB, NKBA = prm.NKBA, NS = prm.NS, NA = prm.NA;
        const bool split = prm.split != 0;
        const int KS1 = prm.KS1, KSL = prm.KSL, PA16 = prm.PA >> 4, KP16 = prm.KP >> 4;
        const uint32_t sB_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | lbo_field;   // descriptor low words, 16-byte units
        const uint32_t stage_u = stage_bytes >> 4, bblk_u = prm.b_block_bytes >> 4, ablk_u = a_blk_bytes >> 4;
        const uint32_t sA_u = smem_u32(sA), a_res = prm.a_res_bytes;
        const uint32_t afull_u = smem_u32(a_full), aempty_u = smem_u32(a_empty);
        const uint32_t bfull_u = smem_u32(b_full), bempty_u = smem_u32(b_empty);
        const uint32_t tfull_u = smem_u32(t_full + iss), tempty_u = smem_u32(t_empty + iss);
        const uint32_t d_tmem = tmem_base + (uint32_t)(iss * G::ACC_STRIDE);
        uint32_t ab = 0, aph = 0, bs = 0, bph = 0, tph = 0;
        int tb = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
          mbar_wait_lean(afull_u + ab * 8, aph);
          tc_fence_after();
          const uint32_t a_lo0 = (((sA_u + ab * a_res) & 0x3FFFFu) >> 4) | lbo_field;
          for (int sweep = 0; sweep < 2; ++sweep) {
            const int nkb = sweep == 0 ? NKBA : NKB;
            const int nseg1 = split ? NKBA : nkb;                       // blocks multiplied with A block kb
            const int lim1 = split ? KS1 : (sweep == 0 ? PA16 : KP16);   // K steps of the first segment
            const bool cross = split && sweep == 1;                     // A.lo x B.first rides on the first segment
            for (int kt = 0; kt < KT; ++kt) {
              const int single_end = cross ? NKBA : 0;                    // blocks [0, single_end): one per stage
              if (tb != iss) {
                // another issuer's tile: watch its stages arrive and acknowledge them (see the generic loop)
                for (int kb = 0; kb < nkb;) {
                  kb += (!split || kb < single_end) ? 1 : min(2, nkb - kb);
                  mbar_wait_lean(bfull_u + bs * 8, bph);
                  if (lane == 0) mbar_arrive(bempty_u + bs * 8);
                  if (++bs == (uint32_t)NS) { bs = 0; bph ^= 1; }
                }
              } else {
                mbar_wait_lean(tempty_u, tph ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < nkb;) {
                  const int nb = (!split || kb < single_end) ? 1 : min(2, nkb - kb);   // blocks of this stage
                  mbar_wait_lean(bfull_u + bs * 8, bph);
                  tc_fence_after();
                  const uint32_t b_lo = sB_lo0 + bs * stage_u;
                  const bool seg1 = kb < nseg1;
                  const int j = seg1 ? kb : kb - nseg1;
                  const int lim = (seg1 ? lim1 : KSL) - j * kpb;        // K steps left in the segment
                  const int n1 = min(nb * kpb, lim);                    // two adjacent blocks: one run of K steps
                  const uint32_t a_lo = a_lo0 + (uint32_t)j * ablk_u;
                  const int n2 = (cross && seg1) ? min(kpb, KSL - kb * kpb) : 0;
                  if (elect_one()) {
                    // K step ks of the run: block ks / kpb of the stage (B blocks bblk_u apart, A blocks ablk_u apart)
#pragma unroll 2
                    for (int ks = 0; ks < n1; ++ks) {   // one K=16 step = two core matrices = 256 bytes = 16 units
                      const uint32_t u = ks >= kpb ? 1u : 0u, kk = (uint32_t)ks - u * (uint32_t)kpb;
                      umma_f16(d_tmem, desc_hi | (a_lo + u * ablk_u + kk * 16u), desc_hi | (b_lo + u * bblk_u + kk * 16u),
                               G::kIdesc, (kb | ks) != 0 ? 1u : 0u);
                    }
                    const uint32_t a2_lo = b_lo + bblk_u;
#pragma unroll 2
                    for (int ks = 0; ks < n2; ++ks)
                      umma_f16(d_tmem, desc_hi | (a2_lo + (uint32_t)ks * 16u), desc_hi | (b_lo + (uint32_t)ks * 16u),
                               G::kIdesc, 1u);
                    umma_commit(bempty_u + bs * 8);
                    if (kb + nb == nkb) umma_commit(tfull_u);
                  }
                  __syncwarp();
                  if (++bs == (uint32_t)NS) { bs = 0; bph ^= 1; }
                  kb += nb;
                }
                tph ^= 1;
              }
              if (++tb == NACC) tb = 0;
            }
          }
          if (elect_one()) umma_commit(aempty_u + ab * 8);   // the A tile may be overwritten once this issuer's MMAs retire
          __syncwarp();
          if (++ab == (uint32_t)NA) { ab = 0; aph ^= 1; }
        }
      }
    } else
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      uint32_t a_base = 0;
      if (prm.NA > 0) {
        mbar_wait<true>(smem_u32(a_full + ab * RS + r), aph);
        a_base = smem_u32(sA + (size_t)(ab * RS + r) * prm.a_res_bytes);
        tc_fence_after();
      }
      for (int sweep = 0; sweep < 2; ++sweep) {
        const int nkb = sweep == 0 ? prm.NKBA : prm.NKB;
        const int kcols = sweep == 0 ? prm.PA : prm.KP;     // operand columns this sweep multiplies
        for (int kt = 0; kt < prm.KT; ++kt) {
          if (RR && tb != iss) {
            // Another issuer's tile: step over its operand blocks -- but WATCH every one of them arrive.  A parity
            // wait is only meaningful within one phase of the barrier; an issuer that jumped ahead (its next tile
            // can be more than a ring lap away) would mistake the previous lap's completion for its own block's.
            // The skipped stage is acknowledged too (b_empty counts every issuer), so the ring can never lap an
            // issuer: every wait below and above stays within one phase of its barrier.
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait<true>(smem_u32(b_full + bs), bph);
              if (lane == 0) mbar_arrive(smem_u32(b_empty + bs));
              if (++bs == prm.NS) { bs = 0; bph ^= 1; }
            }
            if (++tb == NACC) tb = 0;
            continue;
          }
          mbar_wait<true>(smem_u32(t_empty + r * NACC + tb), tph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)((r * NACC + tb) * G::ACC_STRIDE);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait<true>(smem_u32(b_full + bs), bph);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(sB + (size_t)bs * stage_bytes);
            const int ksteps = min(prm.KC, kcols - kb * prm.KC) >> 4;
            if (elect_one()) {
              const uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | lbo_field;
              const uint32_t a_addr = prm.NA > 0 ? a_base + (uint32_t)kb * a_blk_bytes
                                                 : b_addr + prm.b_block_bytes + r * a_blk_bytes;
              const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | lbo_field;
#pragma unroll 4
              for (int ks = 0; ks < ksteps; ++ks)     // one K=16 step = two core matrices = 256 bytes = 16 units
                umma_f16(d_tmem, desc_hi | (a_lo + (uint32_t)ks * 16u), desc_hi | (b_lo + (uint32_t)ks * 16u), G::kIdesc,
                         (kb | ks) != 0 ? 1u : 0u);
              umma_commit(smem_u32(b_empty + bs));       // this row set is done with the stage when its MMAs retire
              if (kb == nkb - 1) umma_commit(smem_u32(t_full + r * NACC + tb));   // accumulator ready
            }
            __syncwarp();
            if (++bs == prm.NS) { bs = 0; bph ^= 1; }
          }
          if (++tb == NACC) { tb = 0; if (!RR) tph ^= 1; }
          if (RR) tph ^= 1;                      // this issuer's own slot: every use flips its phase
        }
      }
      if (prm.NA > 0) {
        if (elect_one()) umma_commit(smem_u32(a_empty + ab * RS + r));   // the A tile may be overwritten
        __syncwarp();
        if (++ab == prm.NA) { ab = 0; aph ^= 1; }
      }
    }
  } else {
    // ================================ epilogue / selection ===========================
    const int ew = warp - 1 - G::NISS;           // epilogue warp index
    const int rset = ew >> 2;                    // row set this warp works on
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int row_t = q * 32 + lane;
    const uint32_t log_base = smem_u32(cand) + (uint32_t)(rset * BM + row_t) * 16;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rset * NACC * G::ACC_STRIDE);
    uint64_t* my_full = t_full + rset * NACC;
    uint64_t* my_empty = t_empty + rset * NACC;
    int ltb = 0, ltph = 0, rtb = 0;              // accumulator ring: load side (slot, phase), release side
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int p = item / prm.QI, qi = item - p * prm.QI;
      const int n = (qi * RS + rset) * BM + row_t;
      const bool row_ok = n < prm.N;
      const int n_c = row_ok ? n : prm.N - 1;
      const float* relrow = HAS_REL ? prm.relpos + (size_t)n_c * prm.M : nullptr;

      // ---- separable bias: A row -> registers, B rows of this warp's 32 rows -> shared memory
      float areg[BIAS > 1 ? KW : 1];
      const float* brow = sepB_s;
      if (BIAS > 1) {
        float* mine = sepB_s + ew * G::SEPW;
        const int first = min(prm.N - 1, (qi * RS + rset) * BM + q * 32);
        const int last = min(prm.N - 1, (qi * RS + rset) * BM + q * 32 + 31);
        const int h0 = first / prm.grid_w;
        const int nh = last / prm.grid_w - h0 + 1;
        __syncwarp();
        // rows are staged zero-padded to sep_mhp = KT * BN / KW entries: no clamping when indexed by key id
        for (int i = lane; i < nh * prm.sep_mhp; i += 32) {
          const int rr = i / prm.sep_mhp, cc = i - rr * prm.sep_mhp;
          mine[i] = cc < prm.sep_mh ? kNegHalfS2 * __ldg(prm.sep_b + (size_t)(h0 + rr) * prm.sep_mh + cc) : 0.f;
        }
        __syncwarp();
        const float* arow = prm.sep_a + (size_t)(n_c % prm.grid_w) * KW;
#pragma unroll
        for (int j = 0; j < KW; ++j) areg[BIAS > 1 ? j : 0] = kNegHalfS2 * __ldg(arow + j);
        brow = mine + (n_c / prm.grid_w - h0) * prm.sep_mhp;
      }

      // ---- chunk pipeline: CH accumulator columns at a time, software pipelined (the TMEM load and
      // the bias terms of chunk i+1 are in flight while chunk i is ranked); shared by both sweeps
      const int total_chunks = prm.KT * NCH;
      auto issue = [&](int ci, Chunk<DENSE, NG>& ch) {
        const int kt = ci / NCH, c = ci - kt * NCH;
        const int m0 = kt * G::BN + c * CH;
        if (c == 0) {
          mbar_wait<false>(smem_u32(my_full + ltb), ltph);
          tc_fence_after();
        }
        tmem_ld36(lane_addr + (uint32_t)(ltb * G::ACC_STRIDE + c * CH), ch.r);
        if (DENSE) {
          if ((prm.M & 3) == 0) {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (m0 + j4 * 4 < prm.M) b4 = __ldg(reinterpret_cast<const float4*>(relrow + m0 + j4 * 4));
              ch.bias[DENSE ? j4 * 4 + 0 : 0] = b4.x; ch.bias[DENSE ? j4 * 4 + 1 : 0] = b4.y;
              ch.bias[DENSE ? j4 * 4 + 2 : 0] = b4.z; ch.bias[DENSE ? j4 * 4 + 3 : 0] = b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) ch.bias[DENSE ? j : 0] = (m0 + j < prm.M) ? __ldg(relrow + m0 + j) : 0.f;
          }
        }
        if (BIAS > 1) {
          const float* brow_c = brow + m0 / KW;
#pragma unroll
          for (int g = 0; g < NG; ++g) ch.bg[g] = brow_c[g];
        } else {
          ch.bg[0] = 0.f;
        }
        if (c == NCH - 1 && ++ltb == NACC) { ltb = 0; ltph ^= 1; }
      };
      auto complete = [&](int ci, Chunk<DENSE, NG>& ch) {
        tmem_ld_wait(ch.r);
        if (ci % NCH == NCH - 1) {                   // whole accumulator is in registers: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(my_empty + rtb));
          if (++rtb == NACC) rtb = 0;
        }
      };
      // score of column j of a chunk WITHOUT the B term of its key group (added per group)
      auto raw = [&](const Chunk<DENSE, NG>& ch, int j) {
        float v = __uint_as_float(ch.r[j]);
        if (DENSE) v = fmaf(ch.bias[DENSE ? j : 0], kNegHalfS2, v);
        else if (BIAS > 1) v += areg[BIAS > 1 ? j % KW : 0];
        return v;
      };
      auto sweep = [&](auto&& process) {
        Chunk<DENSE, NG> c0, c1;
        issue(0, c0);
#pragma unroll 1
        for (int ci = 0; ci < total_chunks; ci += 2) {
          complete(ci, c0);
          if (ci + 1 < total_chunks) issue(ci + 1, c1);
          if (prm.force_rerank != -2) process(ci, c0);
          if (ci + 1 < total_chunks) {
            complete(ci + 1, c1);
            if (ci + 2 < total_chunks) issue(ci + 2, c0);
            if (prm.force_rerank != -2) process(ci + 1, c1);
          }
        }
      };

      // ---- sweep A: TA largest 18-key group maxima (single-plane scores) ------------------------
      float thr_base;
      {
        TopList<TA> la;
        la.init();
        sweep([&](int ci, Chunk<DENSE, NG>& ch) {
          (void)ci;
          float tm[CH / 3];              // triplet maxima, without the B term of the key group
#pragma unroll
          for (int t = 0; t < CH / 3; ++t) tm[t] = fmax3(raw(ch, 3 * t), raw(ch, 3 * t + 1), raw(ch, 3 * t + 2));
          // group size of the list: 18 keys when the row has plenty of groups for its TA entries, finer when
          // the list is long relative to the keys (the excess of keys above the TA-th group maximum grows
          // like TA / (2 * groups)); GA is picked by the host
          if (GA == 18) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float gm;
              if (KW >= 18) {
                gm = fmaxf(fmax3(tm[6 * h], tm[6 * h + 1], tm[6 * h + 2]), fmax3(tm[6 * h + 3], tm[6 * h + 4], tm[6 * h + 5])) +
                     ch.bg[(18 * h) / KW];
              } else {
                gm = fmaxf(fmax3(tm[6 * h], tm[6 * h + 1], tm[6 * h + 2]) + ch.bg[(2 * h) % NG],
                           fmax3(tm[6 * h + 3], tm[6 * h + 4], tm[6 * h + 5]) + ch.bg[(2 * h + 1) % NG]);
              }
              la.insert(gm);
            }
          } else if (GA == 6) {
#pragma unroll
            for (int u = 0; u < CH / 6; ++u)
              la.insert(fmaxf(tm[2 * u] + ch.bg[(6 * u) / KW], tm[2 * u + 1] + ch.bg[(6 * u + 3) / KW]));
          } else {
#pragma unroll
            for (int t = 0; t < CH / 3; ++t) la.insert(tm[t] + ch.bg[(3 * t) / KW]);
          }
        });
        // fewer than TA real groups (tiny M): the floor stays -> everything is logged -> fix-up
        thr_base = la.v[TA - 1] - kSweepSlack;
        // rows past N are all-zero operand rows: every key ties, and logging them would flood the log
        if (!row_ok) thr_base = INFINITY;
      }

      // ---- sweep B: log the triplets that beat the threshold (fp16x3 scores) --------------------
      int cnt = 0;                                 // log entries of this row
      bool overflow = false;
      sweep([&](int ci, Chunk<DENSE, NG>& ch) {
        const int kt = ci / NCH;
        const int m0 = kt * G::BN + (ci - kt * NCH) * CH;
        // id word of a logged triplet: first key id | key group index << 16 (group = slot in brow)
        const int idw = m0 | (BIAS > 1 ? (m0 / KW) << 16 : 0);
        if (DBG && prm.dbg_dist != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (m0 + j < prm.M)
              prm.dbg_dist[((size_t)p * prm.N + n) * prm.M + m0 + j] = (raw(ch, j) + ch.bg[j / KW]) * kScoreToDist;
        }
        // half a chunk (6 triplets) between two capacity checks
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float s[6][3];
          int idv[6];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int j = 3 * (6 * h + u);
#pragma unroll
            for (int i = 0; i < 3; ++i) s[u][i] = raw(ch, j + i);
            idv[u] = idw + (j | (BIAS > 1 ? (j / KW) << 16 : 0));
          }
          bool hit[6];
#pragma unroll
          for (int u = 0; u < 6; ++u)   // the B term of the key group is folded into the threshold
            hit[u] = fmax3(s[u][0], s[u][1], s[u][2]) > thr_base - ch.bg[(3 * (6 * h + u)) / KW];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            // the address is rebuilt from the counter for every store: a running pointer made each increment wait
            // for the previous store to read it (scoreboard), one stall per triplet
            const uint32_t addr = log_base + (uint32_t)cnt * LOG_STRIDE;
            if (hit[u]) {                            // logged without the B term (looked up again from the id word)
              st_shared_v4(addr, s[u][0], s[u][1], s[u][2], __int_as_float(idv[u]));
              ++cnt;
            }
          }
          // The slack slots took the <= 6 triplets of this half chunk.  A row past its LC valid entries had a
          // loose sweep-A threshold (smooth features: its nearest keys share a few 18-key groups): raise the
          // threshold to the T-th largest key logged so far and drop what falls below.  Rare on random data.
          if (__any_sync(0xffffffffu, cnt > LC)) {
            uint32_t lp = log_base + (uint32_t)cnt * LOG_STRIDE;
            log_compact<LOG_STRIDE, T, BIAS>(log_base, lp, thr_base, brow);
            cnt = (int)((lp - log_base) / LOG_STRIDE);
            if (cnt > LC) { overflow = true; cnt = LC; }   // a pile of ties: certify by fix-up
          }
        }
      });

      // ---------------- hand the candidates over -----------------------------------
      // The keys of the logged triplets that reach the threshold go to global
      // memory, slot-major per item so that a warp writes 256 contiguous bytes; knn_finalize_kernel
      // (one thread per row, full occupancy) sorts them, checks the gaps and writes the neighbour ids.
      // Keeping that work here would hold the accumulators -- and the tensor pipe -- for ~15 % of an item.
      {
        const size_t rbase = (size_t)item * G::ROWS + (size_t)(rset * BM + row_t);
        float2* cdst = prm.cand + (size_t)item * prm.cand_slots * G::ROWS + (size_t)(rset * BM + row_t);
        const int mx_cnt = __reduce_max_sync(0xffffffffu, cnt);
        // keys of the logged triplets that reach `tau` -> candidate slots; ADD_BG: the log still holds raw scores
        auto dump = [&](float tau, bool add_bg) {
          int np = 0;
          for (int e = 0; e < mx_cnt; ++e) {
            if (e < cnt) {
              const float4 c = ld_shared_v4(log_base + (uint32_t)e * LOG_STRIDE);
              const int id = (int)(__float_as_uint(c.w) & 0xffffu);
              const float bg = (BIAS > 1 && add_bg) ? brow[__float_as_uint(c.w) >> 16] : 0.f;
              const float sc[3] = {c.x + bg, c.y + bg, c.z + bg};
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                if (sc[i] >= tau && id + i < prm.M) {
                  if (np < prm.cand_slots) cdst[(size_t)np * G::ROWS] = make_float2(sc[i], __int_as_float(id + i));
                  ++np;
                }
              }
            }
          }
          return np;
        };
        // Usually every key above the sweep threshold fits the candidate slots (~T + 1 keys on random
        // features).  Only when some row of the warp has more -- neighbouring keys are often similar -- is the
        // threshold tightened to the T-th largest KEY of the log (which also converts the log to true scores).
        int np = dump(thr_base, true);
        if (__any_sync(0xffffffffu, np > prm.cand_slots)) {
          thr_base = fmaxf(log_tth_key<LOG_STRIDE, T, BIAS>(log_base, cnt, mx_cnt, brow), thr_base);
          np = dump(thr_base, false);
        }
        // every key that is not handed over scores <= thr_base
        // count < 0: the candidate set is not trustworthy (log or slot overflow) -> brute-force fix-up
        prm.cand_count[rbase] = (overflow || np > prm.cand_slots) ? -1 : np;
        prm.cand_thr[rbase] = thr_base;
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
int launch_select(const TcParams& prm, const Plan& pl, int T, int ga, cudaStream_t stream);

}  // namespace tc
}  // namespace gkg

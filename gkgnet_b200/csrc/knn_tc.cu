// Fused pairwise-distance + top-k on the Blackwell tensor cores (tcgen05 / TMEM / TMA bulk).
//
// What it computes (reference: torch_edge.py:39-51, 89-106, 139-149): for every query row n
// of problem p, the k*dilation keys with the smallest
//     dist[n, m] = (|xh_n|^2 - 2 xh_n . yh_m) + |yh_m|^2 + relative_pos[n, m]
// in ascending order, every `dilation`-th kept.  The N x M matrix never leaves the SM:
//
//   * operands: gkg_knn_prepare splits the normalised fp32 rows into fp16 hi/lo parts
//     (scaled by 256 so the lo part stays normal) and writes them K-concatenated,
//         A = [x_hi | x_hi | x_lo | S  S ],   B = [y_hi | y_lo | y_hi | c_hi c_lo]
//     with c = -0.5*|yh|^2*S, so ONE fp16 GEMM with fp32 accumulation yields
//     S^2 (xh.yh - |yh|^2/2) to ~2^-21 relative -- the "fp16x3" split.  Rows are stored
//     in the UMMA no-swizzle K-major core-matrix order, tile by tile, so a whole operand
//     tile is one contiguous TMA bulk copy (cp.async.bulk, mbarrier complete_tx);
//   * warp 0 streams tiles (A: one 128-row query tile, all K; B: 128-key blocks through a
//     ring), warp 1 issues tcgen05.mma (M=128, N=128, K=16 per instruction) into one of
//     four 128-column TMEM accumulators, warps 2-5 drain accumulators with tcgen05.ld
//     (thread == query row), add the bias and keep a per-row candidate list;
//   * selection: threshold filter into a per-thread shared-memory buffer (branch-free),
//     batched insertion into a sorted register list of T = k*d + 2 entries;
//   * exactness: rows whose approximate gaps are below 2*delta are re-ranked with the
//     exact fp32 formula (same arithmetic as knn_exact.cu); rows whose candidate set
//     itself is in doubt go to a tiny exact fix-up kernel.
#include "knn_tc.cuh"

namespace gkg {

namespace {

constexpr int BM = 128;            // query rows per tile  (UMMA M)
constexpr int BN = 144;            // keys per tile        (UMMA N): 4 chunks of 36 = lcm of the
                                   // key-grid widths 9/18/36 of the separable position bias
constexpr int CH = 36;             // columns per epilogue chunk (tcgen05.ld x32 + x4)
constexpr int NTHREADS = 192;      // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int NACC = 3;            // TMEM accumulators
constexpr int ACC_STRIDE = 160;    // TMEM columns between accumulators (3 x 160 <= 512)
constexpr int SEP_B_FLOATS = 2048; // staged rows of the separable bias table B (512 per epilogue warp)
constexpr int CAND_CAP = 40;       // per-row candidate buffer entries (>= max T for the id staging)
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
constexpr size_t kCandBytes = (size_t)BM * CAND_CAP * 8;
constexpr size_t kBarBytes = 1024 + SEP_B_FLOATS * 4;

Plan make_plan(int P, int N, int M, int D) {
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

// ------------------------------------------------------------------------------------
// operand preparation (phase "prepare"): normalise + split + lay out, one pass over the features
// ------------------------------------------------------------------------------------
// A block owns 32 consecutive (padded) rows of one problem.  Phase 1: a warp per row L2-normalises
// the D group channels (F.normalize semantics, torch_edge.py:167-168,173), keeps the fp32 row in
// shared memory and writes it + |xh|^2 for the exact re-rank.  Phase 2: every thread emits 16-byte
// core-matrix rows (8 fp16) of the tensor-core operand:
//   [tile][k-block][row group (tile_rows/8)][k chunk (KC/8)][row (8)][elem (8)]
// so that a warp writes 4 contiguous 128-byte core matrices.
constexpr int PREP_ROWS = 32;

template <typename T, bool IS_KEY>
__global__ void __launch_bounds__(256)
tc_prepare_kernel(const T* __restrict__ feat, int64_t stride_b, int64_t stride_n, float* __restrict__ hat,
                  float* __restrict__ sq, __half* __restrict__ op, int G, int rows, int D, int KP, int KC,
                  int tiles, int tile_rows, int write_hat) {
  extern __shared__ float prep_s[];              // [PREP_ROWS][D] normalised rows + [PREP_ROWS] norms
  float* xs = prep_s;
  float* sqs = prep_s + PREP_ROWS * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long p = blockIdx.y;
  const int g = (int)(p % G);
  const long long b = p / G;
  const int r0 = blockIdx.x * PREP_ROWS;

  for (int rl = warp; rl < PREP_ROWS; rl += 8) {
    const int row = r0 + rl;
    float* dst = xs + rl * D;
    if (row < rows) {
      const T* src = feat + b * stride_b + (long long)row * stride_n + (long long)g * D;
      float ss = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float v = to_f32<T>(src[d]);
        ss = fmaf(v, v, ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float denom = fmaxf(sqrtf(ss), 1e-12f);
      float s2 = 0.f;
      float* gh = hat + (p * rows + row) * (long long)D;
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
  const int total = PREP_ROWS * kc_all;
  for (int ci = threadIdx.x; ci < total; ci += 256) {
    const int r = ci & 7;
    const int kcI = (ci >> 3) % kc_all;
    const int rgl = (ci >> 3) / kc_all;
    const int rl = rgl * 8 + r;
    const int row = r0 + rl;
    const int tile = row / tile_rows;
    if (tile >= tiles) continue;
    const int rg = (row - tile * tile_rows) >> 3;
    const int kb = kcI / kcs, kc = kcI - kb * kcs;
    const bool valid = row < rows;
    const float* src = xs + rl * D;
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
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      float v = 0.f;
      if (c < 3 * D) {
        const int seg = c / D;
        const float x = src[c - seg * D] * kScale;
        const float hi = __half2float(__float2half_rn(x));
        // A = [hi | hi | lo], B = [hi | lo | hi]
        const bool want_lo = IS_KEY ? (seg == 1) : (seg == 2);
        v = want_lo ? (x - hi) : hi;
      } else if (c == 3 * D) {
        v = extra_hi;
      } else if (c == 3 * D + 1) {
        v = extra_lo;
      }
      out[e] = __float2half_rn(v);
    }
    const long long chunk = ((((p * tiles + tile) * nkb + kb) * rgs + rg) * (long long)kcs + kc) * 8 + r;
    *reinterpret_cast<uint4*>(op + chunk * 8) = *reinterpret_cast<const uint4*>(out);
  }
}

// ------------------------------------------------------------------------------------
// sorted candidate list (registers)
// ------------------------------------------------------------------------------------
template <int T>
struct TopList {
  float v[T];
  int id[T];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < T; ++s) { v[s] = INFINITY; id[s] = 0x7fffffff; }
  }
  // requires x < v[T-1]; equal values keep arrival order
  __device__ __forceinline__ void insert(float x, int m) {
#pragma unroll
    for (int s = T - 1; s >= 1; --s) {
      const bool up = x < v[s - 1];
      const bool here = (!up) && (x < v[s]);
      v[s] = up ? v[s - 1] : (here ? x : v[s]);
      id[s] = up ? id[s - 1] : (here ? m : id[s]);
    }
    if (x < v[0]) { v[0] = x; id[0] = m; }
  }
};

// Merge the buffered candidates of every lane into its sorted list (warp-uniform trip count).
// The scan stores (value without the per-key-group bias term, key id); KW > 0 adds brow[id / KW].
template <int T, int KW>
__device__ __forceinline__ void compact_candidates(TopList<T>& top, float& tau, float2*& wp, float2* cbuf,
                                                   const float* brow, int mh_last) {
  const int cnt = (int)(wp - cbuf) / BM;
  const int mx = __reduce_max_sync(0xffffffffu, cnt);
  for (int e = 0; e < mx; ++e) {
    if (e < cnt) {
      const float2 c = cbuf[e * BM];
      const int m = __float_as_int(c.y);
      float v = c.x;
      if (KW > 0) v += brow[min(m / KW, mh_last)];
      if (v < tau) {
        top.insert(v, m);
        tau = top.v[T - 1];
      }
    }
  }
  wp = cbuf;
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(cand) + kCandBytes);
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
    TopList<T> top;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int p = item / prm.QT, qt = item - p * prm.QT;
      const int n = qt * BM + row_t;
      const bool row_ok = n < prm.N;
      const int n_c = row_ok ? n : prm.N - 1;
      const float* relrow = HAS_REL ? prm.relpos + (size_t)n_c * prm.M : nullptr;
      top.init();
      float tau = INFINITY;
      float2* wp = cbuf;                           // next free candidate slot of this row
      constexpr int CKW = BIAS > 1 ? BIAS : 0;

      // ---- separable bias: A row -> registers, B rows of this tile -> shared memory
      float areg[KW];
      const float* brow = sepB_s;
      if (BIAS > 1) {
        // each warp stages the B rows its own 32 query rows need (no cross-warp barrier)
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
                *wp = make_float2(v, __int_as_float(m0 + j));
                wp += BM;
              }
              // the buffer must always have room for the rest of the chunk
              if ((j % 12) == 11 && __any_sync(0xffffffffu, wp > cbuf + (CAND_CAP - 12) * BM)) {
                compact_candidates<T, CKW>(top, tau, wp, cbuf, brow, prm.sep_mh - 1);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(t_empty + tb));
        if (++tb == NACC) { tb = 0; tph ^= 1; }
      }
      compact_candidates<T, CKW>(top, tau, wp, cbuf, brow, prm.sep_mh - 1);

      // ---------------- finalise the row -------------------------------------------
      const int kd = prm.kd;
      bool amb = prm.force_rerank != 0;
#pragma unroll
      for (int s = 0; s + 1 < T; ++s)
        if (s < kd && (top.v[s + 1] - top.v[s]) < 2.f * kDelta) amb = true;
      if (row_ok && amb) {
        // exact fp32 re-rank of the T candidates (identical arithmetic to knn_exact.cu)
        const float* xr = prm.xhat + ((size_t)p * prm.N + n) * prm.D;
        const float xs = prm.xsq[(size_t)p * prm.N + n];
        const float* yb = prm.yhat + (size_t)p * prm.M * prm.D;
        const float* ysb = prm.ysq + (size_t)p * prm.M;
        const float a_last = top.v[T - 1];
        float maxerr = 0.f;
#pragma unroll
        for (int s = 0; s < T; ++s) {
          const int m = top.id[s];
          if (m < prm.M) {
            const float e = exact_dist(xr, yb + (size_t)m * prm.D, prm.D, xs, ysb[m], relrow, m);
            maxerr = fmaxf(maxerr, fabsf((e - xs) - top.v[s]));
            top.v[s] = e;
          } else {
            top.v[s] = INFINITY;
          }
        }
        // odd-even transposition sort by (dist, id)
#pragma unroll
        for (int pass = 0; pass < T; ++pass) {
#pragma unroll
          for (int s = pass & 1; s + 1 < T; s += 2) {
            const bool sw = (top.v[s + 1] < top.v[s]) || (top.v[s + 1] == top.v[s] && top.id[s + 1] < top.id[s]);
            const float tv = sw ? top.v[s] : top.v[s + 1];
            const int ti = sw ? top.id[s] : top.id[s + 1];
            top.v[s] = sw ? top.v[s + 1] : top.v[s];
            top.id[s] = sw ? top.id[s + 1] : top.id[s];
            top.v[s + 1] = tv;
            top.id[s + 1] = ti;
          }
        }
        atomicAdd(prm.stats + 0, 1u);
        atomicMax(prm.stats + 1, __float_as_uint(maxerr));
        // candidate set in doubt?  every non-candidate has approx >= a_last
        float e_kd = -INFINITY;               // sorted ascending: kd-th value == max of the first kd
#pragma unroll
        for (int s = 0; s < T; ++s)
          if (s < kd) e_kd = fmaxf(e_kd, top.v[s]);
        if (a_last - kDelta <= (e_kd - xs) + kDelta && prm.M > T) {
          const int slot = atomicAdd(prm.fix_count, 1);
          prm.fix_rows[slot] = p * prm.N + n;
        }
      }
      // stage the ids in this thread's (now idle) candidate slots so that the dilated pick is a
      // shared-memory index, not a dynamic register index
#pragma unroll
      for (int s = 0; s < T; ++s) cbuf[s * BM].y = __int_as_float(top.id[s]);
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

// ------------------------------------------------------------------------------------
// exact fix-up for rows whose candidate set could not be certified (rare)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_fixup_kernel(const int* __restrict__ count, const int* __restrict__ rows, const float* __restrict__ xhat,
                 const float* __restrict__ xsq, const float* __restrict__ yhat, const float* __restrict__ ysq,
                 const float* __restrict__ relpos, int32_t* __restrict__ idx_out, int N, int M, int D, int k,
                 int dilation) {
  extern __shared__ float dist_s[];            // [M]
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ float sel_v;
  __shared__ int sel_i;
  const int total = *count;
  const int kd = k * dilation;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int id = rows[item];
    const int p = id / N, n = id - p * N;
    const float* xr = xhat + (size_t)id * D;
    const float xs = xsq[id];
    const float* relrow = relpos ? relpos + (size_t)n * M : nullptr;
    for (int m = threadIdx.x; m < M; m += blockDim.x)
      dist_s[m] = exact_dist(xr, yhat + ((size_t)p * M + m) * D, D, xs, ysq[(size_t)p * M + m], relrow, m);
    __syncthreads();
    float last_v = -INFINITY;
    int last_i = -1;
    for (int r = 0; r < kd; ++r) {
      float bv = INFINITY;
      int bi = 0x7fffffff;
      for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float v = dist_s[m];
        const bool after = (v > last_v) || (v == last_v && m > last_i);
        if (after && (v < bv || (v == bv && m < bi))) { bv = v; bi = m; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = bv; red_i[threadIdx.x >> 5] = bi; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
          if (red_v[w] < bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
        sel_v = bv; sel_i = bi;
        if (r % dilation == 0) idx_out[(size_t)id * k + r / dilation] = bi;
      }
      __syncthreads();
      last_v = sel_v;
      last_i = sel_i;
    }
    __syncthreads();
  }
}

struct TcWorkspace {
  __half* a_op; __half* b_op; int* fix_count; int* fix_rows; unsigned int* stats; size_t bytes;
};

TcWorkspace carve_tc(void* base, const Plan& pl, int P, int N) {
  TcWorkspace w;
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t n) { void* p = b ? b + off : nullptr; off += align_up(n, 256); return p; };
  w.a_op = static_cast<__half*>(take(pl.a_op_bytes));
  w.b_op = static_cast<__half*>(take(pl.b_op_bytes));
  w.fix_count = static_cast<int*>(take(256));
  w.stats = reinterpret_cast<unsigned int*>(w.fix_count ? w.fix_count + 4 : nullptr);
  w.fix_rows = static_cast<int*>(take(sizeof(int) * (size_t)P * N));
  w.bytes = off;
  return w;
}

int g_force_rerank = 0;
float* g_dbg_dist = nullptr;
unsigned int g_last_stats[4] = {0, 0, 0, 0};

template <int T, int BIAS>
int launch_select_tb(const TcParams& prm, const Plan& pl, cudaStream_t stream) {
  auto kern = knn_tc_kernel<T, BIAS>;
  size_t smem = pl.smem_bytes < 120 * 1024 ? 120 * 1024 : pl.smem_bytes;   // 512 TMEM columns: 1 CTA / SM
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("knn_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return GKG_ECUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = prm.P * prm.QT;
  const int grid = items < sms ? items : sms;
  kern<<<grid, NTHREADS, smem, stream>>>(prm);
  GKG_CHECK_LAUNCH("knn_tc_kernel");
  return GKG_OK;
}

}  // namespace

bool knn_tc_supported(int N, int M, int D, int k, int dilation) {
  const int kd = k * dilation;
  if (kd + 2 > 38 || kd > M) return false;
  if (N < 1 || M < 1) return false;
  Plan pl = make_plan(1, N, M, D);
  return pl.ok;
}

size_t knn_tc_workspace_bytes(int P, int N, int M, int D, int k, int dilation, bool self_keys) {
  (void)k; (void)dilation; (void)self_keys;
  Plan pl = make_plan(P, N, M, D);
  if (!pl.ok) return 0;
  return carve_tc(nullptr, pl, P, N).bytes;
}

template <typename T>
static int launch_prepare_typed(const KnnWorkspace& w, const TcWorkspace& t, const Plan& pl, const void* x,
                                int64_t x_sb, int64_t x_sn, const void* y, int64_t y_sb, int64_t y_sn, int P,
                                int G, int N, int M, int D, bool self_keys, cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)PREP_ROWS * D + PREP_ROWS);
  {
    dim3 grid((pl.QT * BM + PREP_ROWS - 1) / PREP_ROWS, P);
    tc_prepare_kernel<T, false><<<grid, 256, smem, stream>>>(static_cast<const T*>(x), x_sb, x_sn, w.xhat, w.xsq,
                                                            t.a_op, G, N, D, pl.KP, pl.KC, pl.QT, BM, 1);
    GKG_CHECK_LAUNCH("tc_prepare_kernel<query>");
  }
  {
    dim3 grid((pl.KT * BN + PREP_ROWS - 1) / PREP_ROWS, P);
    const T* src = static_cast<const T*>(self_keys ? x : y);
    tc_prepare_kernel<T, true><<<grid, 256, smem, stream>>>(src, self_keys ? x_sb : y_sb, self_keys ? x_sn : y_sn,
                                                           w.yhat, w.ysq, t.b_op, G, M, D, pl.KP, pl.KC, pl.KT,
                                                           BN, self_keys ? 0 : 1);
    GKG_CHECK_LAUNCH("tc_prepare_kernel<key>");
  }
  return GKG_OK;
}

int launch_knn_tc_prepare(const KnnWorkspace& w, void* extra_ws, const void* x, int64_t x_sb, int64_t x_sn,
                          const void* y, int64_t y_sb, int64_t y_sn, int dtype, int P, int G, int N, int M,
                          int D, bool self_keys, cudaStream_t stream) {
  Plan pl = make_plan(P, N, M, D);
  GKG_CHECK_ARG(pl.ok, "knn_tc: no tiling for D=%d", D);
  GKG_CHECK_ARG(P <= 65535, "knn_tc: B*G=%d > 65535", P);
  TcWorkspace t = carve_tc(extra_ws, pl, P, N);
  if (dtype == GKG_F32)
    return launch_prepare_typed<float>(w, t, pl, x, x_sb, x_sn, y, y_sb, y_sn, P, G, N, M, D, self_keys, stream);
  return launch_prepare_typed<__nv_bfloat16>(w, t, pl, x, x_sb, x_sn, y, y_sb, y_sn, P, G, N, M, D, self_keys,
                                             stream);
}

int launch_knn_tc(const KnnWorkspace& w, void* extra_ws, const float* relpos, const SepBias& sep,
                  int32_t* idx_out, int P, int N, int M, int D, int k, int dilation, bool self_keys,
                  cudaStream_t stream) {
  (void)self_keys;
  Plan pl = make_plan(P, N, M, D);
  GKG_CHECK_ARG(pl.ok, "knn_tc: no tiling for D=%d", D);
  TcWorkspace t = carve_tc(extra_ws, pl, P, N);
  cudaError_t e = cudaMemsetAsync(t.fix_count, 0, 256, stream);
  if (e != cudaSuccess) {
    set_error("knn_tc: memset: %s", cudaGetErrorString(e));
    return GKG_ECUDA;
  }
  count_launch();
  TcParams prm{};
  prm.a_op = t.a_op; prm.b_op = t.b_op;
  prm.xhat = w.xhat; prm.xsq = w.xsq; prm.yhat = w.yhat; prm.ysq = w.ysq;
  prm.relpos = relpos; prm.idx_out = idx_out;
  prm.fix_count = t.fix_count; prm.fix_rows = t.fix_rows; prm.stats = t.stats;
  prm.dbg_dist = g_dbg_dist;
  prm.P = P; prm.N = N; prm.M = M; prm.D = D; prm.k = k; prm.dilation = dilation; prm.kd = k * dilation;
  prm.KP = pl.KP; prm.KC = pl.KC; prm.NKB = pl.NKB; prm.NA = pl.NA; prm.NS = pl.NS; prm.QT = pl.QT; prm.KT = pl.KT;
  prm.a_tile_bytes = pl.a_tile_bytes; prm.b_block_bytes = pl.b_block_bytes;
  prm.force_rerank = g_force_rerank;
  prm.sep_a = sep.a; prm.sep_b = sep.b; prm.grid_w = sep.grid_w > 0 ? sep.grid_w : 1;
  prm.sep_mh = sep.kw > 0 ? M / sep.kw : 1;
  const int T = prm.kd + 2;
  int bias = relpos != nullptr ? 1 : 0;
  if (bias && sep.a != nullptr && sep.b != nullptr && (sep.kw == 9 || sep.kw == 18 || sep.kw == 36) &&
      sep.grid_w > 0 && N % sep.grid_w == 0 && M % sep.kw == 0 &&
      (32 / sep.grid_w + 2) * (M / sep.kw) <= SEP_B_FLOATS / 4)
    bias = sep.kw;
  int rc;
#define GKG_TC_DISPATCH_T(B)                                             \
  (T <= 11 ? launch_select_tb<11, B>(prm, pl, stream)                    \
   : T <= 20 ? launch_select_tb<20, B>(prm, pl, stream)                  \
   : T <= 29 ? launch_select_tb<29, B>(prm, pl, stream)                  \
             : launch_select_tb<38, B>(prm, pl, stream))
  switch (bias) {
    case 0: rc = GKG_TC_DISPATCH_T(0); break;
    case 9: rc = GKG_TC_DISPATCH_T(9); break;
    case 18: rc = GKG_TC_DISPATCH_T(18); break;
    case 36: rc = GKG_TC_DISPATCH_T(36); break;
    default: rc = GKG_TC_DISPATCH_T(1); break;
  }
#undef GKG_TC_DISPATCH_T
  if (rc != GKG_OK) return rc;
  const size_t fsmem = sizeof(float) * (size_t)M;
  GKG_CHECK_ARG(fsmem <= 200 * 1024, "knn_tc: M=%d too large for the fix-up kernel", M);
  static size_t fix_configured = 0;
  if (fsmem > 48 * 1024 && fsmem > fix_configured) {
    cudaFuncSetAttribute(knn_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    fix_configured = fsmem;
  }
  knn_fixup_kernel<<<148, 256, fsmem, stream>>>(t.fix_count, t.fix_rows, w.xhat, w.xsq, w.yhat, w.ysq, relpos,
                                                idx_out, N, M, D, k, dilation);
  GKG_CHECK_LAUNCH("knn_fixup_kernel");
  if (g_dbg_dist != nullptr || g_force_rerank) {   // debug only: expose the counters
    cudaStreamSynchronize(stream);
    cudaMemcpy(g_last_stats, t.fix_count, sizeof(g_last_stats), cudaMemcpyDeviceToHost);
    // layout: [0] fix_count, [4..] stats -> copy stats separately
    unsigned int st[2];
    cudaMemcpy(st, t.stats, sizeof(st), cudaMemcpyDeviceToHost);
    g_last_stats[1] = st[0];
    g_last_stats[2] = st[1];
  }
  return GKG_OK;
}

}  // namespace gkg

// debug hooks (not part of the public header; used by tests through ctypes)
extern "C" void gkg_debug_knn_tc(int force_rerank, float* dbg_dist) {
  gkg::g_force_rerank = force_rerank;
  gkg::g_dbg_dist = dbg_dist;
}
extern "C" void gkg_debug_knn_tc_stats(unsigned int* out3) {
  out3[0] = gkg::g_last_stats[0];
  out3[1] = gkg::g_last_stats[1];
  out3[2] = gkg::g_last_stats[2];
}

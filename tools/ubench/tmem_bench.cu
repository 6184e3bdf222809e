// Micro-benchmarks that decide the kNN kernel design (run on a B200 under gpurun):
//   A  tcgen05.ld throughput with 4 / 8 / 16 warps (x32 / x64 / x128 shapes)
//   B  tcgen05.mma kind::f16 SS-mode issue rate for N = 80 / 144 / 256 (M = 128, K = 16 per instruction)
//   C  mixed operand formats (A = bf16, B = fp16) and f16 accumulators: legal? correct? TMEM packing?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_bench tmem_bench.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000ll) __trap(); }
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

#define LD_X(NREG, STR)                                                                                         \
  template <> __device__ __forceinline__ void tmem_ld<NREG>(uint32_t taddr, uint32_t* r);

template <int NREG> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- A: LDTM throughput
// NW warps; warp w reads lane quarter w % 4; iters loads of NREG columns each, `depth` loads in flight per wait.
template <int NREG, int DEPTH>
__global__ void ldtm_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[DEPTH][NREG];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) tmem_ld<NREG>(tb + (uint32_t)(((i * DEPTH + d) * NREG) & 255), r[d]);
    tmem_wait();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int j = 0; j < NREG; ++j) acc ^= r[d][j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(512) : "memory");
}

template <int NREG, int DEPTH>
void run_ldtm(int nw, long long* d_cyc, uint32_t* d_sink) {
  const int iters = 2000;
  ldtm_kernel<NREG, DEPTH><<<148, nw * 32>>>(iters, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  ldtm_kernel<NREG, DEPTH><<<148, nw * 32>>>(iters, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  long long c[148];
  CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
  const double bytes = (double)iters * DEPTH * NREG * 32 * 4 * nw;
  printf("LDTM x%-3d depth %d warps %2d: %8.0f cycles, %6.1f cyc/load/warp, %7.1f B/clk/SM\n", NREG, DEPTH, nw, avg,
         avg / (iters * DEPTH), bytes / avg);
}

// ---------------------------------------------------------------- B: MMA SS issue rate
// One thread issues `nmma` MMAs (M = 128, N, K = 16 each, rotating over K steps of a smem tile), then commits.
__global__ void mma_rate_kernel(int N, int ksteps, int nmma, uint32_t idesc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero operands (fp16 zeros)
  const int KC = ksteps * 16;
  const int abytes = 128 * KC * 2, bbytes = N * KC * 2;
  for (int i = threadIdx.x; i < (abytes + bbytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + abytes;
    const uint32_t sbo = (uint32_t)(KC / 8) * 128u;
    const long long t0 = clock64();
    const uint64_t ad0 = make_desc(a0, 128, sbo), bd0 = make_desc(b0, 128, sbo);
    int ks = 0; uint32_t slot = 0;
#pragma unroll 4
    for (int i = 0; i < nmma; ++i) {
      umma_f16(tmem_slot + slot, ad0 + (uint64_t)(ks * 16), bd0 + (uint64_t)(ks * 16), idesc, ks != 0);
      if (++ks == ksteps) { ks = 0; slot ^= 256u; }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(512) : "memory");
}

// ---------------------------------------------------------------- C: formats
// One MMA chain: D (128 x N) = A (128 x K) B^T (N x K), operands written by the threads in core-matrix order
// from global arrays (already in the requested 16-bit formats), result dumped from TMEM (raw 32-bit words).
__global__ void mma_fmt_kernel(const uint16_t* A, const uint16_t* B, int N, int K, uint32_t idesc, uint32_t* out, int ncols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  uint16_t* sa = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sb = sa + 128 * K;
  const int sbo_e = (K / 8) * 64;   // elements between 8-row groups
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    sa[(r / 8) * sbo_e + (c / 8) * 64 + (r % 8) * 8 + (c % 8)] = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    sb[(r / 8) * sbo_e + (c / 8) * 64 + (r % 8) * 8 + (c % 8)] = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
    const uint32_t sbo = (uint32_t)(K / 8) * 128u;
    for (int ks = 0; ks < K / 16; ++ks)
      umma_f16(tmem_slot, make_desc(a0 + ks * 256, 128, sbo), make_desc(b0 + ks * 256, 128, sbo), idesc, ks != 0);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t r[16];
      tmem_ld<16>(tmem_slot + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_wait();
      for (int j = 0; j < 16; ++j) out[(size_t)(warp * 32 + lane) * ncols + c0 + j] = r[j];
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(512) : "memory");
}


// ---------------------------------------------------------------- D: MMA + commit round trip
// One thread: (nmma MMAs, commit, wait) repeated `reps` times -> cycles per round trip.
__global__ void mma_roundtrip_kernel(int N, int nmma, int reps, uint32_t idesc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int KC = nmma * 16;
  const int abytes = 128 * KC * 2, bbytes = N * KC * 2;
  for (int i = threadIdx.x; i < (abytes + bbytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + abytes;
    const uint32_t sbo = (uint32_t)(KC / 8) * 128u;
    const uint64_t ad0 = make_desc(a0, 128, sbo), bd0 = make_desc(b0, 128, sbo);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < nmma; ++ks)
        umma_f16(tmem_slot, ad0 + (uint64_t)(ks * 16), bd0 + (uint64_t)(ks * 16), idesc, ks != 0);
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), r & 1);
    }
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(512) : "memory");
}

static uint16_t f2h(float f) { __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static uint16_t f2b(float f) { __nv_bfloat16 h = __float2bfloat16(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static float b2f(uint16_t u) { __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h); }

int main(int argc, char** argv) {
  const int sel_c = argc > 1 ? atoi(argv[1]) : -1;   // >= 0: run only format combo sel_c

  long long* d_cyc; uint32_t* d_sink;
  CK(cudaMalloc(&d_cyc, 148 * 8)); CK(cudaMalloc(&d_sink, 148 * 1024 * 4));
  if (sel_c < 0) {
  printf("== A: tcgen05.ld throughput (148 CTAs, one per SM)\n");
  for (int nw : {4, 8, 16}) {
    run_ldtm<16, 1>(nw, d_cyc, d_sink);
    run_ldtm<16, 2>(nw, d_cyc, d_sink);
    run_ldtm<32, 1>(nw, d_cyc, d_sink);
    run_ldtm<32, 2>(nw, d_cyc, d_sink);
  }
  printf("== D: MMA + commit + wait round trip (one thread), N = 144\n");
  for (int nmma : {1, 3, 6, 13}) {
    const int N = 144;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    CK(cudaFuncSetAttribute(mma_roundtrip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t smem = (size_t)(128 + N) * nmma * 16 * 2;
    mma_roundtrip_kernel<<<148, 128, smem>>>(N, nmma, 500, idesc, d_cyc);
    CK(cudaDeviceSynchronize());
    mma_roundtrip_kernel<<<148, 128, smem>>>(N, nmma, 500, idesc, d_cyc);
    CK(cudaDeviceSynchronize());
    long long c[148];
    CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
    printf("round trip with %2d MMAs: %7.1f cycles\n", nmma, avg / 500);
  }
  printf("== B: tcgen05.mma kind::f16 SS issue rate, M = 128 (one CTA per SM, 148 CTAs)\n");
  for (int N : {80, 144, 160, 256}) {
    for (int ksteps : {3, 13}) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int KC = ksteps * 16;
      const size_t smem = (size_t)(128 + N) * KC * 2;
      CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      const int nmma = ksteps * 400;
      mma_rate_kernel<<<148, 128, smem>>>(N, ksteps, nmma, idesc, d_cyc);
      CK(cudaDeviceSynchronize());
      mma_rate_kernel<<<148, 128, smem>>>(N, ksteps, nmma, idesc, d_cyc);
      CK(cudaDeviceSynchronize());
      long long c[148];
      CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
      printf("MMA N %3d ksteps %2d: %6.1f cycles / MMA (floor %d), smem operand bytes / MMA %d\n", N, ksteps, avg / nmma,
             128 * N / 256, (128 + N) * 32);
    }
  }
  }
  if (sel_c >= 0) {
    const int N = 16, K = 32;
    std::vector<float> a(128 * K), b(N * K);
    srand(1);
    for (auto& v : a) v = (rand() % 2001 - 1000) / 1000.f;
    for (auto& v : b) v = (rand() % 2001 - 1000) / 1000.f;
    uint16_t *dA, *dB; uint32_t* dO;
    CK(cudaMalloc(&dA, 128 * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dO, 128 * 64 * 4));
    // format codes: 0 = f16, 1 = bf16
    for (int cfmt : {1, 0}) for (int af : {0, 1}) for (int bf : {0, 1}) {
      if (sel_c != cfmt * 4 + af * 2 + bf) continue;
      std::vector<uint16_t> ha(128 * K), hb(N * K);
      std::vector<float> ra(128 * K), rb(N * K);
      for (int i = 0; i < 128 * K; ++i) { ha[i] = af ? f2b(a[i]) : f2h(a[i]); ra[i] = af ? b2f(ha[i]) : h2f(ha[i]); }
      for (int i = 0; i < N * K; ++i) { hb[i] = bf ? f2b(b[i]) : f2h(b[i]); rb[i] = bf ? b2f(hb[i]) : h2f(hb[i]); }
      CK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
      CK(cudaMemset(dO, 0xff, 128 * 64 * 4));
      const uint32_t idesc = ((uint32_t)cfmt << 4) | ((uint32_t)af << 7) | ((uint32_t)bf << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      CK(cudaFuncSetAttribute(mma_fmt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      mma_fmt_kernel<<<1, 128, (128 + N) * K * 2>>>(dA, dB, N, K, idesc, dO, 16);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cfmt %d a %d b %d: launch failed: %s\n", cfmt, af, bf, cudaGetErrorString(e)); return 1; }
      std::vector<uint32_t> o(128 * 16);
      CK(cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      if (cfmt == 1) {
        for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
          double ref = 0; for (int k = 0; k < K; ++k) ref += (double)ra[r * K + k] * rb[n * K + k];
          float got; memcpy(&got, &o[r * 16 + n], 4);
          maxerr = fmax(maxerr, fabs(got - ref));
        }
        printf("D f32, A %s, B %s: max |err| vs fp64 of the rounded operands = %.3e\n", af ? "bf16" : "f16", bf ? "bf16" : "f16", maxerr);
      } else {
        // f16 accumulators: print the first words of rows 0, 1 to see the packing
        printf("D f16, A %s, B %s: row 0 words:", af ? "bf16" : "f16", bf ? "bf16" : "f16");
        for (int j = 0; j < 8; ++j) printf(" %08x", o[j]);
        printf("\n   expected row 0:");
        for (int n = 0; n < 8; ++n) { double ref = 0; for (int k = 0; k < K; ++k) ref += (double)ra[k] * rb[n * K + k]; printf(" %04x(%.3f)", f2h((float)ref), ref); }
        printf("\n");
      }
    }
  }
  return 0;
}

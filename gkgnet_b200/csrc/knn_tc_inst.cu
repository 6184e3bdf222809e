// One bias mode of the tcgen05 kNN kernel per translation unit (-DGKG_TC_BIAS=0|1|9|18|36).
#include "knn_tc_kernel.cuh"

namespace gkg {
namespace tc {

template <class G, int T, int BIAS, int GA>
static int launch_select_tb(const TcParams& prm, const Plan& pl, cudaStream_t stream) {
  auto kern = prm.dbg_dist != nullptr ? knn_tc_kernel<G, T, BIAS, GA, true> : knn_tc_kernel<G, T, BIAS, GA, false>;
  const size_t smem = pl.smem_bytes < 120 * 1024 ? 120 * 1024 : pl.smem_bytes;   // 512 TMEM columns: 1 CTA / SM
  static std::atomic<uint64_t> configured[2];
  cudaError_t e = cudaSuccess;
  configure_once_per_device(configured[prm.dbg_dist != nullptr ? 1 : 0], [&] {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
  });
  if (e != cudaSuccess) {
    set_error("knn_tc: cudaFuncSetAttribute(%zu): %s", kSmemBudget, cudaGetErrorString(e));
    return GKG_ECUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = prm.P * prm.QI;
  const int grid = items < sms ? items : sms;
  kern<<<grid, G::NTHREADS, smem, stream>>>(prm);
  GKG_CHECK_LAUNCH("knn_tc_kernel");
  return GKG_OK;
}

// keys per group of the sweep-A list (18 | 6 | 3, see the kernel): GeomA carries all three
template <int T>
static int launch_select_a(const TcParams& prm, const Plan& pl, int ga, cudaStream_t stream) {
  if (ga == 18) return launch_select_tb<GeomA, T, GKG_TC_BIAS, 18>(prm, pl, stream);
  if (ga == 6) return launch_select_tb<GeomA, T, GKG_TC_BIAS, 6>(prm, pl, stream);
  return launch_select_tb<GeomA, T, GKG_TC_BIAS, 3>(prm, pl, stream);
}

template <>
int launch_select<GKG_TC_BIAS>(const TcParams& prm, const Plan& pl, int T, int ga, cudaStream_t stream) {
  if (pl.geom == 1) {   // 256-row items: only planned for lists of <= 20 and 18-key groups
    if (T <= 11) return launch_select_tb<GeomB, 11, GKG_TC_BIAS, 18>(prm, pl, stream);
    return launch_select_tb<GeomB, 20, GKG_TC_BIAS, 18>(prm, pl, stream);
  }
  if (T <= 11) return launch_select_a<11>(prm, pl, ga, stream);
  if (T <= 20) return launch_select_a<20>(prm, pl, ga, stream);
  if (T <= 29) return launch_select_a<29>(prm, pl, ga, stream);
  return launch_select_a<38>(prm, pl, ga, stream);
}

}  // namespace tc
}  // namespace gkg

// Tuned MultiScaleRoIAlign paths (7x7, sampling_ratio 2) — filled in after the generic kernels
// are parity-green; until then every call is declined and the generic kernels run.
#include "roi_common.cuh"

namespace dgod {

int msroi_fwd_fast(const dgod_roi_config*, const RoiDev&, const float*, int, void*, cudaStream_t,
                   int* handled) {
  *handled = 0;
  return DGOD_OK;
}

int msroi_bwd_fast(const dgod_roi_config*, const RoiDev&, const void*, const float*, int,
                   const int32_t*, cudaStream_t, int* handled) {
  *handled = 0;
  return DGOD_OK;
}

}  // namespace dgod

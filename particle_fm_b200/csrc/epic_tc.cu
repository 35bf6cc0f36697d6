// bf16 tcgen05 path -- placeholder until the tensor-core kernel lands.
#include "pfm_internal.cuh"
namespace pfm {
int tc_supported(const pfm_epic* h, int N) {
  (void)h; (void)N;
  set_error("PFM_PREC_BF16: tensor-core path not built yet");
  return PFM_ERR_UNSUPPORTED;
}
int tc_plan_caps(const pfm_epic*, int, int*, int*) { return PFM_ERR_UNSUPPORTED; }
int tc_pack_weights(pfm_epic*, cudaStream_t) { return PFM_ERR_UNSUPPORTED; }
int tc_run(pfm_epic*, const RunArgs&, cudaStream_t) { return PFM_ERR_UNSUPPORTED; }
}  // namespace pfm

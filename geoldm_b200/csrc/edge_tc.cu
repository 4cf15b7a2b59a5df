// tcgen05 / TMEM fused edge kernel (placeholder until the tensor-core path lands).
#include "common.cuh"
namespace geoldm {
int launch_edge_tc(const geoldm_egnn_config& cfg, const geoldm_edge_mlp&, const geoldm_batch&, bool, const float*,
                   const float*, const float*, float*, cudaStream_t) {
  set_error("mma_mode %d: tcgen05 edge kernel not built into this library", cfg.mma_mode);
  return -3;
}
}  // namespace geoldm
extern "C" int geoldm_has_tcgen05(void) { return 0; }

// AdamW(amsgrad) + EMA of the weights as ONE multi-tensor launch (training step, config 5).
//
// Reference: qm9/models.py:169-175 (torch.optim.AdamW(lr, amsgrad=True, weight_decay=1e-12)), utils.py:5-28 (EMA:
// avg = avg * beta + (1 - beta) * new), train_test.py:60-66.  The library's fused optimiser and the two foreach passes of
// the EMA are 24 chunked launches over 278 small tensors (0.7 ms of the captured step, ~330 GB/s); here every 4096-element
// chunk of every tensor is one block of one launch: p, g, m, v, vmax, ema are read and p, m, v, vmax, ema written once.
// The update rule is the library's (decoupled weight decay, bias corrections from the step count, amsgrad maximum kept on
// the un-corrected second moment), computed in fp32 like its fused kernel.
#include "common.cuh"

namespace geoldm {
namespace {
constexpr int OPT_CHUNK = 4096, OPT_T = 256;

__global__ void __launch_bounds__(OPT_T) adamw_ema_kernel(const geoldm_optim_tensor* __restrict__ table,
                                                          const int2* __restrict__ chunk_map, const float* __restrict__ step,
                                                          const float* __restrict__ grad_scale, float lr, float beta1,
                                                          float beta2, float eps, float weight_decay, int amsgrad,
                                                          float ema_beta) {
  const int2 cm = chunk_map[blockIdx.x];                  // {tensor, first element of this chunk}
  const geoldm_optim_tensor T = table[cm.x];
  const float st = *step;                                  // already incremented for this update
  if (T.step && cm.y == 0 && threadIdx.x == 0) *T.step = st;
  const float bc1 = 1.0f - powf(beta1, st), bc2 = 1.0f - powf(beta2, st);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
  const float gs = grad_scale ? *grad_scale : 1.0f;
  const int end = min(T.n, cm.y + OPT_CHUNK);
  for (int i = cm.y + threadIdx.x; i < end; i += OPT_T) {
    float p = T.p[i];
    const float g = T.g[i] * gs;
    float m = T.m[i], v = T.v[i];
    p -= lr * weight_decay * p;
    m = m + (1.0f - beta1) * (g - m);                      // lerp(m, g, 1 - beta1)
    v = beta2 * v + (1.0f - beta2) * g * g;
    float denom;
    if (amsgrad) {
      const float vm = fmaxf(T.vmax[i], v);
      T.vmax[i] = vm;
      denom = sqrtf(vm) / bc2_sqrt + eps;
    } else {
      denom = sqrtf(v) / bc2_sqrt + eps;
    }
    p -= step_size * m / denom;
    T.p[i] = p; T.m[i] = m; T.v[i] = v;
    if (T.ema) T.ema[i] = T.ema[i] * ema_beta + (1.0f - ema_beta) * p;
  }
}
}  // namespace
}  // namespace geoldm

extern "C" int geoldm_adamw_ema_step(const geoldm_optim_tensor* table, const int* chunk_map, int n_chunks, const float* step,
                                     const float* grad_scale, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, int amsgrad, float ema_beta, void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(table != nullptr && chunk_map != nullptr && step != nullptr, "adamw_ema_step: null argument%s", "");
  if (n_chunks == 0) return 0;
  adamw_ema_kernel<<<n_chunks, OPT_T, 0, (cudaStream_t)stream>>>(table, reinterpret_cast<const int2*>(chunk_map), step, grad_scale,
                                                                 lr, beta1, beta2, eps, weight_decay, amsgrad, ema_beta);
  GEOLDM_CHECK_LAUNCH("adamw_ema_kernel");
  return 0;
}
extern "C" int geoldm_optim_chunk(void) { return geoldm::OPT_CHUNK; }

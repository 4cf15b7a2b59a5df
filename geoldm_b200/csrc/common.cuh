// Shared device helpers for the geoldm_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/geoldm_b200.h"

namespace geoldm {

// ---- error plumbing (thread-local text, integer codes) --------------------------------------
void set_error(const char* fmt, ...);
#define GEOLDM_CHECK_LAUNCH(what)                                                      \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      geoldm::set_error("%s: %s", what, cudaGetErrorString(e__));                      \
      return -2;                                                                       \
    }                                                                                  \
  } while (0)
#define GEOLDM_REQUIRE(cond, ...)                                                      \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      geoldm::set_error(__VA_ARGS__);                                                  \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

// ---- math -----------------------------------------------------------------------------------
// SiLU(v) = v / (1 + exp(-v)).  ex2.approx + fast divide: ~2-3 ulp, unbiased; the reference's CPU
// path uses a 1-ulp vectorised expf, so both sit at the fp32 noise floor.
// ex2 / rcp with flush-to-zero: the non-ftz forms add a range check and two predicated multiplies per call to return
// denormal results, which 1 + exp(-v) then rounds away anyway.
__device__ __forceinline__ float ex2_ftz(float v) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float v) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
  return y;
}
__device__ __forceinline__ float silu(float v) {
  return v * rcp_ftz(1.0f + ex2_ftz(v * -1.4426950408889634f));
}
__device__ __forceinline__ float sigmoidf_(float v) { return rcp_ftz(1.0f + ex2_ftz(v * -1.4426950408889634f)); }

// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issue slot) -------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// SiLU of four values held as two packed pairs, in place: 1.5 MUFU operations per value instead of 2.
//   d = 1 + 2^(-v log2 e);  the reciprocals of the lane-wise partners share one rcp:  r = 1 / (dA dB),
//   1/dA = r dB, 1/dB = r dA.   The exponent is clamped at 60 so that dA dB <= 2^121 stays finite; for v < -41.6
//   sigmoid(v) is then 2^-60 instead of e^v (absolute error of SiLU < 1e-16 |v|).  SiLU(0) = 0 exactly; a NaN in one
//   value does not reach its partner (min() drops it before the product).  Error: <= ~3 ulp, like ex2 + rcp per value.
__device__ __forceinline__ void silu_x4(f32x2& A, f32x2& B) {
  const f32x2 nl2e = pk2(-1.4426950408889634f, -1.4426950408889634f);
  const f32x2 one = pk2(1.0f, 1.0f);
  float a0, a1, b0, b1;
  upk2(mul2(A, nl2e), a0, a1);
  upk2(mul2(B, nl2e), b0, b1);
  a0 = ex2_ftz(fminf(a0, 60.0f)); a1 = ex2_ftz(fminf(a1, 60.0f));
  b0 = ex2_ftz(fminf(b0, 60.0f)); b1 = ex2_ftz(fminf(b1, 60.0f));
  const f32x2 dA = add2(pk2(a0, a1), one), dB = add2(pk2(b0, b1), one);
  float p0, p1;
  upk2(mul2(dA, dB), p0, p1);
  const f32x2 r = pk2(rcp_ftz(p0), rcp_ftz(p1));
  A = mul2(A, mul2(r, dB));
  B = mul2(B, mul2(r, dA));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- cp.async (LDGSTS) 16-byte ----------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- per-edge geometry (coord2diff, egnn/egnn_new.py:249-255) ----------------------------------
struct EdgeGeom {
  float r;     // ||x_i - x_j||^2 at block entry
  float d0;    // same on the coordinates at EGNN entry
  float ux, uy, uz;  // (x_i - x_j) / (sqrt(r + 1e-8) + norm_constant)
};
__device__ __forceinline__ EdgeGeom edge_geom(const float* __restrict__ x, const float* __restrict__ x0, int i,
                                              int j, float norm_constant) {
  EdgeGeom g;
  float dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
  // same association as torch.sum((d)**2, 1): ((dx^2 + dy^2) + dz^2), no FMA contraction
  g.r = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  float ex = x0[3 * i] - x0[3 * j], ey = x0[3 * i + 1] - x0[3 * j + 1], ez = x0[3 * i + 2] - x0[3 * j + 2];
  g.d0 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
  float den = __fadd_rn(__fsqrt_rn(__fadd_rn(g.r, 1e-8f)), norm_constant);
  g.ux = __fdiv_rn(dx, den);
  g.uy = __fdiv_rn(dy, den);
  g.uz = __fdiv_rn(dz, den);
  return g;
}

// ---- per-device launch configuration -----------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are PER DEVICE: a process that drives several GPUs
// configures each kernel instantiation once per device ordinal (small fixed table; ordinals past it reconfigure every time).
constexpr int GEOLDM_MAX_DEVICES = 32;
struct DeviceOnce {
  bool done[GEOLDM_MAX_DEVICES] = {};
  int sm_count[GEOLDM_MAX_DEVICES] = {};
};
// returns the current device ordinal clamped into the table, sets `fresh` when this (kernel, device) pair is new
inline int device_slot(DeviceOnce& st, bool& fresh) {
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = (dev >= 0 && dev < GEOLDM_MAX_DEVICES) ? dev : GEOLDM_MAX_DEVICES - 1;
  fresh = !st.done[slot] || slot != dev;
  if (fresh) cudaDeviceGetAttribute(&st.sm_count[slot], cudaDevAttrMultiProcessorCount, dev);
  return slot;
}

// ---- launchers implemented in the .cu files ----------------------------------------------------
int launch_edge_simt(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                     const float* pq, const float* x, const float* x0, float* out, cudaStream_t st);
int launch_edge_tc(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                   const float* pq, int pq_ld, const float* x, const float* x0, float* out, cudaStream_t st);
int launch_linear(const float* a1, int k1, const float* a2, int k2, float a2_div, const float* wt,
                  const float* bias, const float* res, int epi, float* out, int m, int n, cudaStream_t st);

int launch_linear_tc(int H, int terms, const float* a1, int k1, const float* a2, int k2, float a2_div,
                     const void* w_pack, int n_blocks, const float* bias, const float* res, int epi, float* out, int m,
                     cudaStream_t st, float* zero_buf = nullptr);
int launch_tc_selftest(int H, int terms, const float* pq, const int* edge_i, const int* tile_row, int n_tile,
                       int n_rows, const void* w_pack, float* out, cudaStream_t st);
// fp16-split variants (edge_tc16.cu), selected by terms == 16 / GEOLDM_MMA_3XF16
int launch_edge_tc16(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                     const float* pq, int pq_ld, const float* x, const float* x0, const float* r_edge,
                     const float* d0_edge, const float* u_edge, float* out, cudaStream_t st);
// out[e] = |x_i - x_j|^2 for every packed edge (same association as edge_geom); u_out (optional, [E][4]) receives the
// normalised difference (x_i - x_j) / (sqrt(r + 1e-8) + norm_constant) of coord2diff (egnn_new.py:249-255)
int launch_edge_dist(const geoldm_batch& b, const float* x, float* out, float* u_out, float norm_constant,
                     cudaStream_t st);
int launch_linear_tc16(int H, const float* a1, int k1, const float* a2, int k2, float a2_div, const void* w_pack,
                       int n_blocks, const float* bias, const float* res, int epi, float* out, int m, cudaStream_t st,
                       float* zero_buf = nullptr);   // zero_buf: [m][n_blocks H] buffer cleared by the residual epilogue
// fused node chain of one GCL + the following first-layer projections (edge_tc16.cu: chain16_kernel)
int launch_node_chain16(int H, const float* h, const float* agg, float agg_div, const void* pack1, const float* b1,
                        const void* pack2, const float* b2, const void* pack3, const float* b3, int n_pb, float* h_out,
                        float* pq_out, float* zero_buf, int m, cudaStream_t st);
int launch_linear_tc16_grad(int H, const float* a_rows, int ld, const void* w_pack, float* out, int m, unsigned* amax_scratch,
                            int amax_ready, cudaStream_t st);
int launch_tc16_selftest(int H, const float* pq, const int* edge_i, const int* tile_row, int n_tile, int n_rows,
                         const void* w_pack, float* out, cudaStream_t st);
int launch_embed(int n_node, int H, int F, const float* h_in, const float* w, const float* b, float* h,
                 cudaStream_t st);
int launch_outproj(int n_node, int H, int Fo, const float* h, const float* w, const float* b, float* out,
                   cudaStream_t st);
int launch_coord_update(int n3, const float* x0, const float* dx, float* xagg, float div, float* dx_next,
                        float* x_next, cudaStream_t st, bool zero_xagg = false);

}  // namespace geoldm

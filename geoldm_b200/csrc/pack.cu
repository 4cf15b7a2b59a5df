// Tensor-core operand packing of weight matrices (see geoldm_tc_pack in include/geoldm_b200.h).
#include "common.cuh"

namespace geoldm {
// round-to-nearest-even to tf32 (10 explicit mantissa bits), result has the low 13 bits cleared
__device__ __forceinline__ float tf32_rne(float v) {
  uint32_t b = __float_as_uint(v);
  b = (b + 0xFFFu + ((b >> 13) & 1u)) & 0xFFFFE000u;
  return __uint_as_float(b);
}

// one thread per weight element: W[n][k] -> block n/H, slab k/32, hi and lo images
__global__ void tc_pack_kernel(int H, const float* __restrict__ w, int n_out, int k, float* __restrict__ pack) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_out * k) return;
  const int n = (int)(idx / k), kk = (int)(idx % k);
  const int nb = n / H, nl = n % H, slab = kk / 32, kl = kk % 32;
  const int n_slabs = k / 32;
  const float v = w[idx];
  const float hi = tf32_rne(v);
  const float lo = tf32_rne(v - hi);
  // stage = (slab, N-half): [hi image | lo image], each H/2 rows x 32 floats in canonical SWIZZLE_128B K-major order
  const int NH = H / 2, half = nl / NH, rl = nl % NH;
  const int off = (rl >> 3) * 256 + (rl & 7) * 32 + ((((kl >> 2) ^ (rl & 7)) << 2) | (kl & 3));
  float* img = pack + ((size_t)((nb * n_slabs + slab) * 2 + half) * 2) * (size_t)NH * 32;
  img[off] = hi;
  img[(size_t)NH * 32 + off] = lo;
}
}  // namespace geoldm

extern "C" {
size_t geoldm_tc_pack_bytes(int H, int n_out, int k) { return (size_t)n_out * k * 2 * sizeof(float) + 0 * H; }

int geoldm_tc_pack(int H, const float* w, int n_out, int k, void* pack, void* stream) {
  GEOLDM_REQUIRE(H % 32 == 0 && H <= 256 && n_out % H == 0 && k % 64 == 0, "tc_pack: H=%d n_out=%d k=%d", H, n_out, k);
  const size_t tot = (size_t)n_out * k;
  if (tot == 0) return 0;
  geoldm::tc_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      H, w, n_out, k, reinterpret_cast<float*>(pack));
  GEOLDM_CHECK_LAUNCH("tc_pack_kernel");
  return 0;
}
}

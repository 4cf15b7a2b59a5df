// Training-side GEMM:  C[N][K] += A[M][N]^T * B[M][K]   (weight gradients dW = dY^T X of every Linear layer).
// fp32 FFMA, 64 x 64 output tile per CTA, the reduction over M is split across blockIdx.z and combined with
// atomicAdd into the (caller-zeroed) output.  Together with geoldm_linear (forward and dX = dY W) this gives the
// autograd path of geoldm_b200/train.py its three GEMMs per Linear (reference: nn.Linear backward through ATen addmm).
#include "common.cuh"

namespace geoldm {
namespace {
constexpr int TN_BN = 64, TN_BK = 64, TN_BM = 16, TN_T = 256;

__global__ void __launch_bounds__(TN_T) gemm_tn_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                      int ldb, float* __restrict__ c, int ldc, int m, int n, int k,
                                                      int m_per_split) {
  __shared__ __align__(16) float As[2][TN_BM][TN_BN];
  __shared__ __align__(16) float Bs[2][TN_BM][TN_BK];
  const int t = threadIdx.x;
  const int n0 = blockIdx.y * TN_BN, k0 = blockIdx.x * TN_BK;
  const int m_begin = blockIdx.z * m_per_split;
  const int m_end = min(m, m_begin + m_per_split);
  const int tr = t >> 4, tc = t & 15;            // 16 x 16 threads, 4 x 4 outputs each
  // loaders: thread -> (row = t / 16, 4 consecutive columns)
  const int lr = t >> 4, lc = (t & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  auto load = [&](int mrow, int buf) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    const int mm = mrow + lr;
    if (mm < m_end) {
      const float* pa = a + (size_t)mm * lda + n0 + lc;
      const float* pb = b + (size_t)mm * ldb + k0 + lc;
      if (n0 + lc + 3 < n) va = *reinterpret_cast<const float4*>(pa);
      else { float tmp[4] = {0.f, 0.f, 0.f, 0.f}; for (int e = 0; e < 4; ++e) if (n0 + lc + e < n) tmp[e] = pa[e]; va = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]); }
      if (k0 + lc + 3 < k) vb = *reinterpret_cast<const float4*>(pb);
      else { float tmp[4] = {0.f, 0.f, 0.f, 0.f}; for (int e = 0; e < 4; ++e) if (k0 + lc + e < k) tmp[e] = pb[e]; vb = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]); }
    }
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = va;
    *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = vb;
  };

  int buf = 0;
  if (m_begin < m_end) load(m_begin, 0);
  __syncthreads();
  for (int mrow = m_begin; mrow < m_end; mrow += TN_BM) {
    if (mrow + TN_BM < m_end) load(mrow + TN_BM, buf ^ 1);
#pragma unroll
    for (int mm = 0; mm < TN_BM; ++mm) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][mm][tr * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][mm][tc * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nn = n0 + tr * 4 + i;
    if (nn >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + tc * 4 + j;
      if (kk < k) atomicAdd(c + (size_t)nn * ldc + kk, acc[i][j]);
    }
  }
}
}  // namespace
}  // namespace geoldm

extern "C" int geoldm_gemm_tn(const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n, int k,
                              void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "gemm_tn: lda=%d ldb=%d must be multiples of 4", lda, ldb);
  if (m == 0 || n == 0 || k == 0) return 0;
  const int tiles = ((n + TN_BN - 1) / TN_BN) * ((k + TN_BK - 1) / TN_BK);
  int splits = (4 * 148 + tiles - 1) / tiles;                  // aim at ~4 CTAs per SM
  const int max_splits = (m + 4 * TN_BM - 1) / (4 * TN_BM);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per_split = ((m + splits - 1) / splits + TN_BM - 1) / TN_BM * TN_BM;
  splits = (m + m_per_split - 1) / m_per_split;
  dim3 grid((k + TN_BK - 1) / TN_BK, (n + TN_BN - 1) / TN_BN, splits);
  gemm_tn_kernel<<<grid, TN_T, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, c, ldc, m, n, k, m_per_split);
  GEOLDM_CHECK_LAUNCH("gemm_tn_kernel");
  return 0;
}

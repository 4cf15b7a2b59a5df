// Node-level fp32 GEMM with fused prologue/epilogue  —  GEOLDM_MMA_FP32_SIMT.
//
//   out[M][N] = epi( a1[M][k1] * wt[0:k1][N] + (a2[M][k2] / a2_div) * wt[k1:k1+k2][N] + bias[N] )
//
// Used for (reference: egnn/egnn_new.py)
//   * the split first edge layer  P|Q = h * [W1[:, :H]^T | W1[:, H:2H]^T] + [b1|0]        (:14-16, :77-78)
//   * node_mlp.0 on cat[h, agg] without materialising the cat, + SiLU                     (:20-22, :53-54)
//   * node_mlp.2 + residual                                                               (:23, :55)
// 128 x 64 tile, 16-wide k-slabs, 8 x 4 register tile per thread, cp.async double buffering for the
// weight slab, register prefetch + transposing store for the activation slab.
#include "common.cuh"

namespace geoldm {
namespace {
constexpr int BM = 128, BN = 64, BK = 16, NT = 256, AST = BM + 4;

struct LinArgs {
  const float* a1; const float* a2; const float* wt; const float* bias; const float* res; float* out;
  int k1, k2, m, n; float a2_div;
};

template <int EPI>
__global__ void __launch_bounds__(NT, 2) linear_kernel(const LinArgs a) {
  __shared__ __align__(16) float As[2][BK][AST];
  __shared__ __align__(16) float Ws[2][BK][BN];
  const int t = threadIdx.x;
  const int rg = t >> 4, cg = t & 15;            // 16 x 16 thread grid
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const int K = a.k1 + a.k2, NS = K / BK;

  // activation loader: thread -> (row = t/2, 8 consecutive k)
  const int lrow = t >> 1, lk = (t & 1) * 8;
  const bool lvalid = row0 + lrow < a.m;
  // weight loader: thread -> (k = t/16, 4 consecutive columns)
  const int wk = t >> 4, wc4 = (t & 15) * 4;
  const bool wvalid = col0 + wc4 < a.n;

  // accumulators as packed fp32 pairs along the output columns (fma.rn.f32x2: two IEEE FMAs per issue slot; the plain
  // three-register FFMA issues every second cycle per scheduler on sm_100), same products and summation order
  f32x2 acc2[8][2];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc2[r][0] = acc2[r][1] = pk2(0.f, 0.f);

  float4 pf0, pf1;
  auto a_load = [&](int s) {
    pf0 = make_float4(0.f, 0.f, 0.f, 0.f); pf1 = pf0;
    if (lvalid) {
      const int k0 = s * BK + lk;
      const float* src = (k0 < a.k1) ? a.a1 + (size_t)(row0 + lrow) * a.k1 + k0
                                     : a.a2 + (size_t)(row0 + lrow) * a.k2 + (k0 - a.k1);
      pf0 = __ldg(reinterpret_cast<const float4*>(src));
      pf1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      if (k0 >= a.k1 && a.a2_div != 1.0f) {
        const float d = a.a2_div;
        pf0.x = __fdiv_rn(pf0.x, d); pf0.y = __fdiv_rn(pf0.y, d); pf0.z = __fdiv_rn(pf0.z, d); pf0.w = __fdiv_rn(pf0.w, d);
        pf1.x = __fdiv_rn(pf1.x, d); pf1.y = __fdiv_rn(pf1.y, d); pf1.z = __fdiv_rn(pf1.z, d); pf1.w = __fdiv_rn(pf1.w, d);
      }
    }
  };
  auto a_store = [&](int buf) {
    float v[8] = {pf0.x, pf0.y, pf0.z, pf0.w, pf1.x, pf1.y, pf1.z, pf1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) As[buf][lk + e][lrow] = v[e];
  };
  auto w_load = [&](int s, int buf) {
    if (wvalid)
      cp_async16(&Ws[buf][wk][wc4], a.wt + (size_t)(s * BK + wk) * a.n + col0 + wc4);
    else
      *reinterpret_cast<float4*>(&Ws[buf][wk][wc4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    cp_async_commit();
  };

  w_load(0, 0);
  a_load(0);
  a_store(0);
  for (int s = 0; s < NS; ++s) {
    const int buf = s & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (s + 1 < NS) { w_load(s + 1, buf ^ 1); a_load(s + 1); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][rg * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + rg * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Ws[buf][k][cg * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const f32x2 b01 = pk2(b.x, b.y), b23 = pk2(b.z, b.w);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const f32x2 aa = pk2(av[r], av[r]);
        acc2[r][0] = fma2(aa, b01, acc2[r][0]);
        acc2[r][1] = fma2(aa, b23, acc2[r][1]);
      }
    }
    if (s + 1 < NS) a_store(buf ^ 1);
  }

  float acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r) { upk2(acc2[r][0], acc[r][0], acc[r][1]); upk2(acc2[r][1], acc[r][2], acc[r][3]); }
  const int col = col0 + cg * 4;
  if (col < a.n) {
    float4 bias = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = row0 + (r >> 2) * 64 + rg * 4 + (r & 3);
      if (row < a.m) {
        float4 o = make_float4(acc[r][0] + bias.x, acc[r][1] + bias.y, acc[r][2] + bias.z, acc[r][3] + bias.w);
        if (EPI == 1) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
        if (EPI == 2) {
          float4 rs = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)row * a.n + col));
          o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
        }
        *reinterpret_cast<float4*>(a.out + (size_t)row * a.n + col) = o;
      }
    }
  }
}

// Same contraction for node-level row counts (a few thousand rows): 32 x 64 tiles, 128 threads, 4 x 4 outputs per thread.
// At m = 1158 (64 QM9 molecules) the 128 x 64 tiles above are 30 CTAs on 148 SMs, each walking the whole K loop: 18 us per
// launch, 144 launches per training step; four times as many, four times shorter CTAs bring that to the latency of the K
// loop itself.  That loop is a chain of (L2 load -> shared memory -> barrier) per 16-wide slab with next to no arithmetic to
// hide it behind, so both operands arrive through a 4-stage cp.async ring: three slabs are in flight while one is consumed
// (A in its global row-major form: a thread reads four k values of a row as one 16-byte vector).
constexpr int SBM = 32, SNT = 128, SST = 4, SAK = BK + 4;
#ifndef GEOLDM_LS_MINB
#define GEOLDM_LS_MINB 4   // less than one CTA per SM at node-level row counts: registers are free (8 -> 4: 9.7 -> 9.1 us per launch)
#endif
template <int EPI>
__global__ void __launch_bounds__(SNT, GEOLDM_LS_MINB) linear_small_kernel(const LinArgs a) {
  __shared__ __align__(16) float As[SST][SBM][SAK];
  __shared__ __align__(16) float Ws[SST][BK][BN];
  const int t = threadIdx.x;
  const int rg = t >> 4, cg = t & 15;            // 8 x 16 thread grid, 4 x 4 outputs each
  const int row0 = blockIdx.y * SBM, col0 = blockIdx.x * BN;
  const int K = a.k1 + a.k2, NS = K / BK;
  const int lrow = t >> 2, lk = (t & 3) * 4;     // activation loader: (row, 4 consecutive k) = one 16-byte copy per slab
  const int grow = min(row0 + lrow, a.m - 1);    // rows past the end: a copy of the last row, computed and never stored
  const int wk = t >> 4, wc4 = (t & 15) * 4;     // weight loader: rows wk and wk + 8 of the slab, 4 consecutive columns
  const bool wvalid = col0 + wc4 < a.n;
  f32x2 acc2[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) acc2[r][0] = acc2[r][1] = pk2(0.f, 0.f);
  auto load = [&](int s, int st) {
    const int k0 = s * BK + lk;
    const float* src = (k0 < a.k1) ? a.a1 + (size_t)grow * a.k1 + k0 : a.a2 + (size_t)grow * a.k2 + (k0 - a.k1);
    cp_async16(&As[st][lrow][lk], src);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = wk + 8 * h;
      if (wvalid) cp_async16(&Ws[st][kk][wc4], a.wt + (size_t)(s * BK + kk) * a.n + col0 + wc4);
      else *reinterpret_cast<float4*>(&Ws[st][kk][wc4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
#pragma unroll
  for (int s = 0; s < SST - 1; ++s) {            // one commit group per slab, empty groups past the end keep the count uniform
    if (s < NS) load(s, s);
    cp_async_commit();
  }
  for (int s = 0; s < NS; ++s) {
    cp_async_wait<SST - 2>();                    // slab s has landed (this thread's copies; the barrier covers the others)
    __syncthreads();                             // ... and every thread is done with the stage refilled below (slab s - 1)
    if (s + SST - 1 < NS) load(s + SST - 1, (s + SST - 1) % SST);
    cp_async_commit();
    const int st = s % SST;
    const bool div = s * BK >= a.k1 && a.a2_div != 1.0f;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      float av[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(&As[st][rg * 4 + r][k4]);
        av[r][0] = v.x; av[r][1] = v.y; av[r][2] = v.z; av[r][3] = v.w;
        if (div) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) av[r][kk] = __fdiv_rn(av[r][kk], a.a2_div);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b = *reinterpret_cast<const float4*>(&Ws[st][k4 + kk][cg * 4]);
        const f32x2 b01 = pk2(b.x, b.y), b23 = pk2(b.z, b.w);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const f32x2 aa = pk2(av[r][kk], av[r][kk]);
          acc2[r][0] = fma2(aa, b01, acc2[r][0]);
          acc2[r][1] = fma2(aa, b23, acc2[r][1]);
        }
      }
    }
  }
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { upk2(acc2[r][0], acc[r][0], acc[r][1]); upk2(acc2[r][1], acc[r][2], acc[r][3]); }
  const int col = col0 + cg * 4;
  if (col < a.n) {
    const float4 bias = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = row0 + rg * 4 + r;
      if (row < a.m) {
        float4 o = make_float4(acc[r][0] + bias.x, acc[r][1] + bias.y, acc[r][2] + bias.z, acc[r][3] + bias.w);
        if (EPI == 1) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
        if (EPI == 2) {
          const float4 rs = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)row * a.n + col));
          o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
        }
        *reinterpret_cast<float4*>(a.out + (size_t)row * a.n + col) = o;
      }
    }
  }
}
}  // namespace

int launch_linear(const float* a1, int k1, const float* a2, int k2, float a2_div, const float* wt,
                  const float* bias, const float* res, int epi, float* out, int m, int n, cudaStream_t st) {
  GEOLDM_REQUIRE(k1 % BK == 0 && k2 % BK == 0 && (k1 + k2) > 0, "linear: k1=%d k2=%d must be multiples of %d", k1, k2, BK);
  GEOLDM_REQUIRE(n % 4 == 0, "linear: n=%d must be a multiple of 4", n);
  GEOLDM_REQUIRE(epi != 2 || res != nullptr, "linear: residual epilogue needs res");
  if (m == 0) return 0;
  LinArgs a{a1, a2, wt, bias, res, out, k1, k2, m, n, a2_div};
#ifndef GEOLDM_LINEAR_SMALL_M
#define GEOLDM_LINEAR_SMALL_M 8192
#endif
  if (m <= GEOLDM_LINEAR_SMALL_M) {   // node-level row counts: small tiles fill the SMs
    dim3 sgrid((n + BN - 1) / BN, (m + SBM - 1) / SBM);
    switch (epi) {
      case 0: linear_small_kernel<0><<<sgrid, SNT, 0, st>>>(a); break;
      case 1: linear_small_kernel<1><<<sgrid, SNT, 0, st>>>(a); break;
      case 2: linear_small_kernel<2><<<sgrid, SNT, 0, st>>>(a); break;
      default: set_error("linear: bad epilogue %d", epi); return -1;
    }
    GEOLDM_CHECK_LAUNCH("linear_small_kernel");
    return 0;
  }
  dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM);
  switch (epi) {
    case 0: linear_kernel<0><<<grid, NT, 0, st>>>(a); break;
    case 1: linear_kernel<1><<<grid, NT, 0, st>>>(a); break;
    case 2: linear_kernel<2><<<grid, NT, 0, st>>>(a); break;
    default: set_error("linear: bad epilogue %d", epi); return -1;
  }
  GEOLDM_CHECK_LAUNCH("linear_kernel");
  return 0;
}
}  // namespace geoldm

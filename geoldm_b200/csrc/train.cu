// Training-side GEMM:  C[N][K] += A[M][N]^T * B[M][K]   (weight gradients dW = dY^T X of every Linear layer).
// fp32 FFMA, 64 x 64 output tile per CTA, the reduction over M is split across blockIdx.z and combined with
// atomicAdd into the (caller-zeroed) output.  Together with geoldm_linear (forward and dX = dY W) this gives the
// autograd path of geoldm_b200/train.py its three GEMMs per Linear (reference: nn.Linear backward through ATen addmm).
#include "common.cuh"

namespace geoldm {
namespace {
#ifndef GEOLDM_TN_RT
#define GEOLDM_TN_RT 8
#endif
// TN_RT x 4 outputs per thread: 4 -> 256 threads, 8 -> 128 threads per 64 x 64 tile (fewer shared-memory operand reads per FMA)
constexpr int TN_BN = 64, TN_BK = 64, TN_BM = 16, TN_RT = GEOLDM_TN_RT, TN_T = 16 * (TN_BN / TN_RT);
constexpr int TN_LR = TN_T / 16, TN_LPT = TN_BM / TN_LR;      // loader rows per pass, passes per slab

__global__ void __launch_bounds__(TN_T) gemm_tn_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b,
                                                      int ldb, float* __restrict__ c, int ldc, int m, int n, int k,
                                                      int m_per_split, float* __restrict__ colsum) {
  __shared__ __align__(16) float As[2][TN_BM][TN_BN];
  __shared__ __align__(16) float Bs[2][TN_BM][TN_BK];
  const int t = threadIdx.x;
  const int n0 = blockIdx.y * TN_BN, k0 = blockIdx.x * TN_BK;
  const int m_begin = blockIdx.z * m_per_split;
  const int m_end = min(m, m_begin + m_per_split);
  const int tr = t >> 4, tc = t & 15;            // (64 / TN_RT) x 16 threads, TN_RT x 4 outputs each
  // loaders: thread -> (row = t / 16, 4 consecutive columns)
  const int lr = t >> 4, lc = (t & 15) * 4;
  // accumulators as packed fp32 pairs along the output columns: fma.rn.f32x2 performs two IEEE FMAs per issue slot (the
  // plain three-register FFMA issues every second cycle per scheduler on sm_100), same products and summation order
  f32x2 acc2[TN_RT][2];
#pragma unroll
  for (int i = 0; i < TN_RT; ++i) acc2[i][0] = acc2[i][1] = pk2(0.f, 0.f);

  // global -> registers (issued before the FMAs of the current slab) -> shared (after them): the L2 latency of the next
  // 16-row slab is hidden behind 256 FMAs per thread instead of being exposed at every slab
  float4 va[TN_LPT], vb[TN_LPT];
  // bias gradient (colsum[n] += sum_m A[m][n], optional): the k0 == 0 blocks add up the A values they stage anyway, one
  // float4 per thread and slab, and combine the row phases through shared memory at the end
  const bool do_colsum = colsum != nullptr && blockIdx.x == 0;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  auto g_load = [&](int mrow) {
#pragma unroll
    for (int q = 0; q < TN_LPT; ++q) {
      va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q];
      const int mm = mrow + lr + q * TN_LR;
      if (mm < m_end) {
        const float* pa = a + (size_t)mm * lda + n0 + lc;
        const float* pb = b + (size_t)mm * ldb + k0 + lc;
        if (n0 + lc + 3 < n) va[q] = __ldg(reinterpret_cast<const float4*>(pa));
        else { float tmp[4] = {0.f, 0.f, 0.f, 0.f}; for (int e = 0; e < 4; ++e) if (n0 + lc + e < n) tmp[e] = pa[e]; va[q] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]); }
        if (k0 + lc + 3 < k) vb[q] = __ldg(reinterpret_cast<const float4*>(pb));
        else { float tmp[4] = {0.f, 0.f, 0.f, 0.f}; for (int e = 0; e < 4; ++e) if (k0 + lc + e < k) tmp[e] = pb[e]; vb[q] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]); }
      }
    }
  };
  auto s_store = [&](int buf) {
#pragma unroll
    for (int q = 0; q < TN_LPT; ++q) {
      cs.x += va[q].x; cs.y += va[q].y; cs.z += va[q].z; cs.w += va[q].w;
      *reinterpret_cast<float4*>(&As[buf][lr + q * TN_LR][lc]) = va[q];
      *reinterpret_cast<float4*>(&Bs[buf][lr + q * TN_LR][lc]) = vb[q];
    }
  };

  int buf = 0;
  if (m_begin < m_end) { g_load(m_begin); s_store(0); }
  __syncthreads();
  for (int mrow = m_begin; mrow < m_end; mrow += TN_BM) {
    const bool more = mrow + TN_BM < m_end;
    if (more) g_load(mrow + TN_BM);
#pragma unroll
    for (int mm = 0; mm < TN_BM; ++mm) {
      float ar[TN_RT];
#pragma unroll
      for (int h = 0; h < TN_RT / 4; ++h) {
        const float4 av = *reinterpret_cast<const float4*>(&As[buf][mm][tr * TN_RT + 4 * h]);
        ar[4 * h] = av.x; ar[4 * h + 1] = av.y; ar[4 * h + 2] = av.z; ar[4 * h + 3] = av.w;
      }
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][mm][tc * 4]);
      const f32x2 b01 = pk2(bv.x, bv.y), b23 = pk2(bv.z, bv.w);
#pragma unroll
      for (int i = 0; i < TN_RT; ++i) {
        const f32x2 aa = pk2(ar[i], ar[i]);
        acc2[i][0] = fma2(aa, b01, acc2[i][0]);
        acc2[i][1] = fma2(aa, b23, acc2[i][1]);
      }
    }
    if (more) s_store(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const bool vec_out = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(c) & 15u) == 0;
  float acc[TN_RT][4];
#pragma unroll
  for (int i = 0; i < TN_RT; ++i) { upk2(acc2[i][0], acc[i][0], acc[i][1]); upk2(acc2[i][1], acc[i][2], acc[i][3]); }
#pragma unroll
  for (int i = 0; i < TN_RT; ++i) {
    const int nn = n0 + tr * TN_RT + i;
    if (nn >= n) continue;
    const int kk0 = k0 + tc * 4;
    if (vec_out && kk0 + 3 < k) {     // one 16-byte vector reduction instead of four scalar ones (L2 reduction operations / 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + (size_t)nn * ldc + kk0), "f"(acc[i][0]),
                   "f"(acc[i][1]), "f"(acc[i][2]), "f"(acc[i][3]) : "memory");
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = kk0 + j;
      if (kk < k) atomicAdd(c + (size_t)nn * ldc + kk, acc[i][j]);
    }
  }
  if (do_colsum) {                                   // block-uniform
    __syncthreads();                                 // (the loop ends with a barrier; kept explicit for the reuse of As)
    *reinterpret_cast<float4*>(&As[0][lr][lc]) = cs;
    __syncthreads();
    if (t < TN_BN && n0 + t < n) {
      float v = 0.f;
#pragma unroll
      for (int rr = 0; rr < TN_LR; ++rr) v += As[0][rr][t];
      atomicAdd(colsum + n0 + t, v);
    }
  }
}
}  // namespace
}  // namespace geoldm

// ---------------------------------------------------------------------------------------------------------------
// Fused element-wise stages of the edge MLPs for the autograd path (train.py): one warp per edge, lanes stride over the
// H columns (coalesced 128-byte rows), everything recomputed from the GEMM inputs/outputs in the backward pass so that
// no [E, H] intermediate besides the GEMM operands is kept.  ex2.approx sigmoid with a correctly rounded division.
//   act   : a[e][k] = SiLU(P[i_e][k] + Q[j_e][k] + r_e w_r[k] + d0_e w_d[k])                  (egnn_new.py:30-36, split form)
//   tail  : m = SiLU(mpre + b2);  gate: agg[i_e] += m sigmoid(m.w_att + b_att) / div        (:37-44, :258-267)
//           head: sc[e] = m.w                                                                 (:86-90)
// ---------------------------------------------------------------------------------------------------------------
namespace geoldm {
namespace {
// lane -> columns: 16-byte groups, lane owns columns 128 g + 4 lane .. + 3 of group g (H <= 256, H % 4 == 0): every row
// access is one float4 per lane (a warp covers 512 contiguous bytes) and the scatter-adds are 16-byte vector reductions
// (red.global.add.v4.f32: a quarter of the L2 reduction operations of the scalar form, which bounded edge_act_bwd)
constexpr int MAXG = 2;

__device__ __forceinline__ void ld4(const float* p, float (&o)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void st4(float* p, const float (&o)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void red4(float* p, const float (&o)[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
}

// ex2.approx (2 ulp) + correctly rounded division: with rcp.approx as well the worst config-5 gradient tensor moved from
// 6e-6 to 1.03e-5 of the fp64 reference run (north_star's tolerance is 1e-5), with accurate expf the kernels are 10 % slower
#ifndef GEOLDM_TRAIN_SIGM
#define GEOLDM_TRAIN_SIGM 2
#endif
__device__ __forceinline__ float sigm(float v) {
#if GEOLDM_TRAIN_SIGM == 0
  return 1.0f / (1.0f + expf(-v));
#elif GEOLDM_TRAIN_SIGM == 1
  return sigmoidf_(v);
#else
  return __fdiv_rn(1.0f, 1.0f + ex2_ftz(v * -1.4426950408889634f));
#endif
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-column partial sums of the 8 warps of a block -> shared memory -> ONE atomicAdd per column and block.  Every warp
// adding its own partials to the same [H] global vector serialises in the L2 atomic unit (thousands of reductions onto
// the same few 128-byte lines): measured 80-93 us per launch before, mostly that.
constexpr int TRAIN_WARPS = 8;
// backward kernels: few fat blocks (every block ends with one reduction per column into the weight gradients), all resident
// at once: 3 blocks of 256 threads per SM at <= 80 registers (4 per SM spill, 2 per SM leave half of the warp slots empty)
#ifndef GEOLDM_TRAIN_BWD_BLOCKS_PER_SM
#define GEOLDM_TRAIN_BWD_BLOCKS_PER_SM 3
#endif
__device__ __forceinline__ void block_column_add(float (*red)[128 * MAXG], const float (&acc)[MAXG][4], int H, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < MAXG; ++g) st4(&red[w][128 * g + 4 * lane], acc[g]);
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int ww = 0; ww < TRAIN_WARPS; ++ww) v += red[ww][c];
    atomicAdd(out + c, v);
  }
  __syncthreads();
}

__global__ void edge_act_fwd_kernel(int E, int H, const float* __restrict__ pq, int pq_ld, const float* __restrict__ r,
                                    const float* __restrict__ d0, const float* __restrict__ w_rd,
                                    const int* __restrict__ ei, const int* __restrict__ ej, float* __restrict__ a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  float wr[MAXG][4], wd[MAXG][4];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    const int c0 = 128 * g + 4 * lane;
    if (c0 < H) { ld4(w_rd + c0, wr[g]); ld4(w_rd + H + c0, wd[g]); }
  }
  for (int e = warp; e < E; e += nwarp) {
    const float* P = pq + (size_t)ei[e] * pq_ld;
    const float* Q = pq + (size_t)ej[e] * pq_ld + H;
    const float re = r[e], de = d0[e];
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
      if (c0 < H) {
        float p[4], q[4], o[4];
        ld4(P + c0, p); ld4(Q + c0, q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float z = p[t] + q[t] + re * wr[g][t] + de * wd[g][t];
          o[t] = z * sigm(z);
        }
        st4(a + (size_t)e * H + c0, o);
      }
    }
  }
}

__global__ void __launch_bounds__(32 * TRAIN_WARPS, GEOLDM_TRAIN_BWD_BLOCKS_PER_SM) edge_act_bwd_kernel(int E, int H, const float* __restrict__ pq, int pq_ld, const float* __restrict__ r,
                                    const float* __restrict__ d0, const float* __restrict__ w_rd,
                                    const int* __restrict__ ei, const int* __restrict__ ej, const float* __restrict__ da,
                                    float* __restrict__ dpq, float* __restrict__ dr, float* __restrict__ dd0,
                                    float* __restrict__ dw_rd) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  float acc_r[MAXG][4], acc_d[MAXG][4], acc_p[MAXG][4], wr[MAXG][4], wd[MAXG][4];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
#pragma unroll
    for (int t = 0; t < 4; ++t) acc_r[g][t] = acc_d[g][t] = acc_p[g][t] = wr[g][t] = wd[g][t] = 0.f;
    const int c0 = 128 * g + 4 * lane;
    if (c0 < H) { ld4(w_rd + c0, wr[g]); ld4(w_rd + H + c0, wd[g]); }
  }
  // a warp owns a CONTIGUOUS range of edge rows: rows are sorted by receiver, so the receiver-side gradient dP_i is summed
  // in registers over the run (fixed order) and leaves as one vector reduction per (run piece, 4 columns) instead of one
  // per edge; the sender-side gradient dQ_j is one vector reduction per (edge, 4 columns)
  const int per = (E + nwarp - 1) / nwarp;
  const int e_begin = min(E, warp * per), e_end = min(E, e_begin + per);
  int cur_i = -1;
  auto flush_p = [&]() {
    if (cur_i < 0) return;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
      if (c0 < H) red4(dpq + (size_t)cur_i * pq_ld + c0, acc_p[g]);
#pragma unroll
      for (int t = 0; t < 4; ++t) acc_p[g][t] = 0.f;
    }
  };
  for (int e = e_begin; e < e_end; ++e) {
    const int i = ei[e], j = ej[e];
    if (i != cur_i) { flush_p(); cur_i = i; }
    const float* P = pq + (size_t)i * pq_ld;
    const float* Q = pq + (size_t)j * pq_ld + H;
    const float re = r[e], de = d0[e];
    float sr = 0.f, sd = 0.f;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
      if (c0 < H) {
        float p[4], q[4], u[4], gq[4];
        ld4(P + c0, p); ld4(Q + c0, q); ld4(da + (size_t)e * H + c0, u);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float z = p[t] + q[t] + re * wr[g][t] + de * wd[g][t];
          const float s = sigm(z);
          const float gv = u[t] * (s * (1.0f + z * (1.0f - s)));     // d SiLU / dz
          gq[t] = gv;
          acc_p[g][t] += gv;
          sr += gv * wr[g][t]; sd += gv * wd[g][t];
          acc_r[g][t] += gv * re; acc_d[g][t] += gv * de;
        }
        red4(dpq + (size_t)j * pq_ld + H + c0, gq);
      }
    }
    sr = warp_sum_f(sr); sd = warp_sum_f(sd);
    if (lane == 0) { dr[e] = sr; dd0[e] = sd; }
  }
  flush_p();
  __shared__ __align__(16) float red[TRAIN_WARPS][128 * MAXG];
  block_column_add(red, acc_r, H, dw_rd);
  block_column_add(red, acc_d, H, dw_rd + H);
}

// gate != 0: agg[i] += m * sigmoid(m.w + b) / div (attention != 0) or m / div; gate == 0: sc[e] = m.w
__global__ void edge_tail_fwd_kernel(int E, int H, const float* __restrict__ mpre, const float* __restrict__ b2,
                                     const float* __restrict__ w, const float* __restrict__ bw, int gate, int attention,
                                     const int* __restrict__ ei, float inv_div, float* __restrict__ agg,
                                     float* __restrict__ sc) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  // contiguous edge range per warp: the messages of a receiver run are summed in registers (fixed order), one vector
  // reduction per (run piece, 4 columns)
  const int per = (E + nwarp - 1) / nwarp;
  const int e_begin = min(E, warp * per), e_end = min(E, e_begin + per);
  int cur_i = -1;
  float acc_a[MAXG][4], bb[MAXG][4], ww[MAXG][4];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
#pragma unroll
    for (int t = 0; t < 4; ++t) acc_a[g][t] = bb[g][t] = ww[g][t] = 0.f;
    const int c0 = 128 * g + 4 * lane;
    if (c0 < H) { ld4(b2 + c0, bb[g]); if (w) ld4(w + c0, ww[g]); }
  }
  auto flush_a = [&]() {
    if (cur_i < 0) return;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
      if (c0 < H) red4(agg + (size_t)cur_i * H + c0, acc_a[g]);
#pragma unroll
      for (int t = 0; t < 4; ++t) acc_a[g][t] = 0.f;
    }
  };
  for (int e = e_begin; e < e_end; ++e) {
    float m[MAXG][4];
    float dot = 0.f;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
#pragma unroll
      for (int t = 0; t < 4; ++t) m[g][t] = 0.f;
      if (c0 < H) {
        float v[4];
        ld4(mpre + (size_t)e * H + c0, v);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float z = v[t] + bb[g][t];
          m[g][t] = z * sigm(z);
          dot += m[g][t] * ww[g][t];
        }
      }
    }
    dot = warp_sum_f(dot);
    if (!gate) { if (lane == 0) sc[e] = dot; continue; }
    const float gt = attention ? sigm(dot + bw[0]) : 1.0f;
    const int i = ei[e];
    if (i != cur_i) { flush_a(); cur_i = i; }
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc_a[g][t] += m[g][t] * gt * inv_div;
  }
  if (gate) flush_a();
}

__global__ void __launch_bounds__(32 * TRAIN_WARPS, GEOLDM_TRAIN_BWD_BLOCKS_PER_SM) edge_tail_bwd_kernel(int E, int H, const float* __restrict__ mpre, const float* __restrict__ b2,
                                     const float* __restrict__ w, const float* __restrict__ bw, int gate, int attention,
                                     const int* __restrict__ ei, float inv_div, const float* __restrict__ dagg,
                                     const float* __restrict__ dsc, float* __restrict__ dmpre, float* __restrict__ db2,
                                     float* __restrict__ dw, float* __restrict__ dbw, double* __restrict__ bw_scratch,
                                     unsigned* __restrict__ damax) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  float acc_b[MAXG][4], acc_w[MAXG][4], bb[MAXG][4], ww[MAXG][4];
  double acc_bw = 0.0;      // attention-bias gradient: a cancelling sum over ALL edges, kept in double and summed in a fixed order
  float omax = 0.f;         // max |dmpre| of this thread (-> damax: the scale of the input-gradient GEMM that consumes dmpre)
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
#pragma unroll
    for (int t = 0; t < 4; ++t) acc_b[g][t] = acc_w[g][t] = bb[g][t] = ww[g][t] = 0.f;
    const int c0 = 128 * g + 4 * lane;
    if (c0 < H) { ld4(b2 + c0, bb[g]); if (w) ld4(w + c0, ww[g]); }
  }
  for (int e = warp; e < E; e += nwarp) {
    float m[MAXG][4], dz[MAXG][4], de[MAXG][4];
    float dot = 0.f, dg = 0.f;
    const float* din = gate ? dagg + (size_t)ei[e] * H : nullptr;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
#pragma unroll
      for (int t = 0; t < 4; ++t) m[g][t] = dz[g][t] = de[g][t] = 0.f;
      if (c0 < H) {
        float v[4], d4[4] = {0.f, 0.f, 0.f, 0.f};
        ld4(mpre + (size_t)e * H + c0, v);
        if (gate) ld4(din + c0, d4);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float z = v[t] + bb[g][t];
          const float s = sigm(z);
          m[g][t] = z * s;
          dz[g][t] = s * (1.0f + z * (1.0f - s));
          dot += m[g][t] * ww[g][t];
          if (gate) { de[g][t] = d4[t] * inv_div; dg += de[g][t] * m[g][t]; }
        }
      }
    }
    dot = warp_sum_f(dot);
    float gt = 1.0f, ds;
    if (gate) {
      dg = warp_sum_f(dg);
      ds = 0.f;
      if (attention) { gt = sigm(dot + bw[0]); ds = dg * gt * (1.0f - gt); }
    } else {
      ds = dsc[e];
    }
    if (lane == 0 && dbw) acc_bw += (double)ds;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      const int c0 = 128 * g + 4 * lane;
      if (c0 < H) {
        float o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float dm = gate ? de[g][t] * gt : 0.f;
          if (w) { dm += ds * ww[g][t]; acc_w[g][t] += ds * m[g][t]; }
          o[t] = dm * dz[g][t];
          acc_b[g][t] += o[t];
          omax = fmaxf(omax, fabsf(o[t]));
        }
        st4(dmpre + (size_t)e * H + c0, o);
      }
    }
  }
  __shared__ __align__(16) float red[TRAIN_WARPS][128 * MAXG];
  block_column_add(red, acc_b, H, db2);
  if (dw) block_column_add(red, acc_w, H, dw);
  if (damax) {                                    // one atomic per block (non-negative floats order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o));
    if (lane == 0) red[0][threadIdx.x >> 5] = omax;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = red[0][0];
#pragma unroll
      for (int k = 1; k < TRAIN_WARPS; ++k) m = fmaxf(m, red[0][k]);
      atomicMax(damax, __float_as_uint(m));
    }
    __syncthreads();
  }
  if (dbw) {
    // block partial (double, fixed warp order) -> scratch[block]; the LAST block to finish adds all partials in a fixed
    // order (strided per thread, then a shared-memory tree: the whole block takes part, a single thread walking 592
    // dependent L2 loads cost tens of microseconds per launch) and accumulates the total into dbw: deterministic and
    // accurate to double rounding.  (Float atomics in arrival order made this 5e-6-sized scalar wander by up to 3e-5
    // relative from run to run.)
    __shared__ double red_bw[32 * TRAIN_WARPS];
    __shared__ bool is_last;
    if (lane == 0) red_bw[threadIdx.x >> 5] = acc_bw;
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < TRAIN_WARPS; ++k) v += red_bw[k];
      if (bw_scratch) {
        bw_scratch[blockIdx.x] = v;
        __threadfence();
        unsigned* counter = reinterpret_cast<unsigned*>(bw_scratch + gridDim.x);
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
      } else {
        atomicAdd(dbw, (float)v);
        is_last = false;
      }
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      double v = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) v += reinterpret_cast<volatile double*>(bw_scratch)[b];
      red_bw[threadIdx.x] = v;
      __syncthreads();
      for (int s = (32 * TRAIN_WARPS) / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red_bw[threadIdx.x] += red_bw[threadIdx.x + s];
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        dbw[0] += (float)red_bw[0];
        *reinterpret_cast<unsigned*>(bw_scratch + gridDim.x) = 0u;        // the scratch is reusable
      }
    }
  }
}

// ---- coord2diff and the coordinate update of EquivariantUpdate for the autograd path (one thread per edge) ------------
//   r[e] = |x_i - x_j|^2,  u[e] = (x_i - x_j) / (sqrt(r + 1e-8) + c)                         (egnn_new.py:249-255)
//   step[i] += u[e] * (tanh(sc[e]) * range | sc[e]) / div                                     (egnn_new.py:91-99)
// One launch each instead of the 8 + 8 library launches of the same arithmetic (and 14 + 12 in their backward passes).
__global__ void coord2diff_fwd_kernel(int E, const float* __restrict__ x, const int* __restrict__ ei,
                                      const int* __restrict__ ej, float c, float* __restrict__ r, float* __restrict__ u) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = ei[e], j = ej[e];
  const float dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
  const float rr = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));   // torch.sum(d ** 2, 1)
  r[e] = rr;
  if (u) {
    const float den = __fadd_rn(__fsqrt_rn(__fadd_rn(rr, 1e-8f)), c);
    u[3 * e] = __fdiv_rn(dx, den); u[3 * e + 1] = __fdiv_rn(dy, den); u[3 * e + 2] = __fdiv_rn(dz, den);
  }
}
// gx[i] += g, gx[j] -= g with g = gu / den - (gu . d) d / (den^2 n) + 2 gr d,  n = sqrt(r + 1e-8), den = n + c
__global__ void coord2diff_bwd_kernel(int E, const float* __restrict__ x, const int* __restrict__ ei,
                                      const int* __restrict__ ej, float c, const float* __restrict__ gr,
                                      const float* __restrict__ gu, float* __restrict__ gx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = ei[e], j = ej[e];
  const float d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
  float g[3] = {0.f, 0.f, 0.f};
  if (gu) {
    const float rr = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    const float n = sqrtf(rr + 1e-8f), den = n + c;
    const float g0 = gu[3 * e], g1 = gu[3 * e + 1], g2 = gu[3 * e + 2];
    const float a = (g0 * d[0] + g1 * d[1] + g2 * d[2]) / (den * den * n);
    g[0] = g0 / den - a * d[0]; g[1] = g1 / den - a * d[1]; g[2] = g2 / den - a * d[2];
  }
  if (gr) {
    const float t = 2.0f * gr[e];
    g[0] += t * d[0]; g[1] += t * d[1]; g[2] += t * d[2];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { atomicAdd(gx + 3 * i + k, g[k]); atomicAdd(gx + 3 * j + k, -g[k]); }
}
__global__ void coord_step_fwd_kernel(int E, const float* __restrict__ u, const float* __restrict__ sc,
                                      const int* __restrict__ ei, int use_tanh, float range, float div,
                                      float* __restrict__ step) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = ei[e];
  const float s = sc[e];
  const float th = use_tanh ? tanhf(s) : s;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float t = u[3 * e + k] * th;                 // (coord_diff * tanh(phi)) * coords_range, then / normalization_factor
    if (use_tanh) t *= range;
    atomicAdd(step + 3 * i + k, __fdiv_rn(t, div));
  }
}
__global__ void coord_step_bwd_kernel(int E, const float* __restrict__ u, const float* __restrict__ sc,
                                      const int* __restrict__ ei, int use_tanh, float range, float div,
                                      const float* __restrict__ gstep, float* __restrict__ gu, float* __restrict__ gsc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = ei[e];
  const float s = sc[e];
  const float th = use_tanh ? tanhf(s) : s;
  const float f = use_tanh ? th * range : th;                     // d trans / d u
  const float dth = use_tanh ? (1.0f - th * th) * range : 1.0f;   // d (tanh(s) range) / d s
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float gt = gstep[3 * i + k] / div;
    gu[3 * e + k] = gt * f;
    dot += gt * u[3 * e + k];
  }
  gsc[e] = dot * dth;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int train_grid(int E) { int b = (E + 7) / 8; return b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b); }
inline int train_grid_fat(int E, int per_sm) {
  int b = (E + 7) / 8;
  return b < 1 ? 1 : (b > 148 * per_sm ? 148 * per_sm : b);
}
inline int train_grid_bwd(int E) { return train_grid_fat(E, GEOLDM_TRAIN_BWD_BLOCKS_PER_SM); }
}  // namespace
}  // namespace geoldm

extern "C" {
int geoldm_train_edge_act_fwd(int n_edge, int H, const float* pq, int pq_ld, const float* r, const float* d0,
                              const float* w_rd, const int* edge_i, const int* edge_j, float* a, void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(H > 0 && H <= 128 * MAXG && H % 4 == 0, "train_edge_act: H=%d not a multiple of 4 in (0, %d]", H, 128 * MAXG);
  GEOLDM_REQUIRE(pq_ld % 4 == 0 && al16(pq) && al16(w_rd) && al16(a), "train_edge_act_fwd: pq / w_rd / a must be 16-byte aligned, pq_ld=%d a multiple of 4", pq_ld);
  if (n_edge == 0) return 0;
  edge_act_fwd_kernel<<<train_grid(n_edge), 256, 0, (cudaStream_t)stream>>>(n_edge, H, pq, pq_ld, r, d0, w_rd, edge_i, edge_j, a);
  GEOLDM_CHECK_LAUNCH("edge_act_fwd_kernel");
  return 0;
}
int geoldm_train_edge_act_bwd(int n_edge, int H, const float* pq, int pq_ld, const float* r, const float* d0,
                              const float* w_rd, const int* edge_i, const int* edge_j, const float* da, float* dpq,
                              float* dr, float* dd0, float* dw_rd, void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(H > 0 && H <= 128 * MAXG && H % 4 == 0, "train_edge_act: H=%d not a multiple of 4 in (0, %d]", H, 128 * MAXG);
  GEOLDM_REQUIRE(pq_ld % 4 == 0 && al16(pq) && al16(w_rd) && al16(da) && al16(dpq), "train_edge_act_bwd: pq / w_rd / da / dpq must be 16-byte aligned, pq_ld=%d a multiple of 4", pq_ld);
  if (n_edge == 0) return 0;
  edge_act_bwd_kernel<<<train_grid_bwd(n_edge), 32 * TRAIN_WARPS, 0, (cudaStream_t)stream>>>(n_edge, H, pq, pq_ld, r, d0, w_rd, edge_i, edge_j, da,
                                                                           dpq, dr, dd0, dw_rd);
  GEOLDM_CHECK_LAUNCH("edge_act_bwd_kernel");
  return 0;
}
int geoldm_train_edge_tail_fwd(int n_edge, int H, const float* mpre, const float* b2, const float* w, const float* bw,
                               int gate, int attention, const int* edge_i, float div, float* agg, float* sc,
                               void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(H > 0 && H <= 128 * MAXG && H % 4 == 0, "train_edge_tail: H=%d not a multiple of 4 in (0, %d]", H, 128 * MAXG);
  GEOLDM_REQUIRE(gate ? (agg != nullptr && (!attention || (w && bw))) : (sc != nullptr && w != nullptr), "train_edge_tail: bad arguments%s", "");
  GEOLDM_REQUIRE(al16(mpre) && al16(b2) && al16(w) && al16(agg), "train_edge_tail_fwd: mpre / b2 / w / agg must be 16-byte aligned%s", "");
  if (n_edge == 0) return 0;
  edge_tail_fwd_kernel<<<train_grid_fat(n_edge, 4), 256, 0, (cudaStream_t)stream>>>(n_edge, H, mpre, b2, attention || !gate ? w : nullptr, bw, gate,
                                                                            attention, edge_i, 1.0f / div, agg, sc);
  GEOLDM_CHECK_LAUNCH("edge_tail_fwd_kernel");
  return 0;
}
int geoldm_train_edge_tail_bwd(int n_edge, int H, const float* mpre, const float* b2, const float* w, const float* bw,
                               int gate, int attention, const int* edge_i, float div, const float* dagg, const float* dsc,
                               float* dmpre, float* db2, float* dw, float* dbw, double* dbw_scratch, void* damax,
                               void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(H > 0 && H <= 128 * MAXG && H % 4 == 0, "train_edge_tail: H=%d not a multiple of 4 in (0, %d]", H, 128 * MAXG);
  GEOLDM_REQUIRE(al16(mpre) && al16(b2) && al16(w) && al16(dagg) && al16(dmpre), "train_edge_tail_bwd: mpre / b2 / w / dagg / dmpre must be 16-byte aligned%s", "");
  if (n_edge == 0) return 0;
  const bool use_w = attention || !gate;
  edge_tail_bwd_kernel<<<train_grid_bwd(n_edge), 32 * TRAIN_WARPS, 0, (cudaStream_t)stream>>>(n_edge, H, mpre, b2, use_w ? w : nullptr, bw, gate, attention,
                                                                            edge_i, 1.0f / div, dagg, dsc, dmpre, db2,
                                                                            use_w ? dw : nullptr, (gate && attention) ? dbw : nullptr,
                                                                            dbw_scratch, reinterpret_cast<unsigned*>(damax));
  GEOLDM_CHECK_LAUNCH("edge_tail_bwd_kernel");
  return 0;
}
}

extern "C" {
int geoldm_train_coord2diff_fwd(int n_edge, const float* x, const int* edge_i, const int* edge_j, float norm_constant,
                                float* r, float* u, void* stream) {
  using namespace geoldm;
  if (n_edge == 0) return 0;
  coord2diff_fwd_kernel<<<(n_edge + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_edge, x, edge_i, edge_j, norm_constant, r, u);
  GEOLDM_CHECK_LAUNCH("coord2diff_fwd_kernel");
  return 0;
}
int geoldm_train_coord2diff_bwd(int n_edge, const float* x, const int* edge_i, const int* edge_j, float norm_constant,
                                const float* gr, const float* gu, float* gx, void* stream) {
  using namespace geoldm;
  if (n_edge == 0 || (gr == nullptr && gu == nullptr)) return 0;
  coord2diff_bwd_kernel<<<(n_edge + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_edge, x, edge_i, edge_j, norm_constant, gr, gu, gx);
  GEOLDM_CHECK_LAUNCH("coord2diff_bwd_kernel");
  return 0;
}
int geoldm_train_coord_step_fwd(int n_edge, const float* u, const float* sc, const int* edge_i, int use_tanh,
                                float coords_range, float div, float* step, void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(div != 0.f, "train_coord_step: div must be non-zero%s", "");
  if (n_edge == 0) return 0;
  coord_step_fwd_kernel<<<(n_edge + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_edge, u, sc, edge_i, use_tanh, coords_range, div, step);
  GEOLDM_CHECK_LAUNCH("coord_step_fwd_kernel");
  return 0;
}
int geoldm_train_coord_step_bwd(int n_edge, const float* u, const float* sc, const int* edge_i, int use_tanh,
                                float coords_range, float div, const float* gstep, float* gu, float* gsc, void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(div != 0.f, "train_coord_step: div must be non-zero%s", "");
  if (n_edge == 0) return 0;
  coord_step_bwd_kernel<<<(n_edge + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_edge, u, sc, edge_i, use_tanh, coords_range, div, gstep, gu, gsc);
  GEOLDM_CHECK_LAUNCH("coord_step_bwd_kernel");
  return 0;
}
}

extern "C" int geoldm_train_bwd_blocks(int n_edge) { return geoldm::train_grid_bwd(n_edge); }

extern "C" int geoldm_gemm_tn_bias(const float* a, int lda, const float* b, int ldb, float* c, int ldc, float* colsum,
                                   int m, int n, int k, void* stream) {
  using namespace geoldm;
  GEOLDM_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "gemm_tn: lda=%d ldb=%d must be multiples of 4", lda, ldb);
  if (m == 0 || n == 0 || k == 0) return 0;
  const int tiles = ((n + TN_BN - 1) / TN_BN) * ((k + TN_BK - 1) / TN_BK);
#ifndef GEOLDM_TN_CTAS_PER_SM
#define GEOLDM_TN_CTAS_PER_SM 3
#endif
  int splits = (GEOLDM_TN_CTAS_PER_SM * (256 / TN_T) * 148) / tiles;          // one resident wave: 3 CTAs of 66 registers x 256 threads per SM
  const int max_splits = (m + 4 * TN_BM - 1) / (4 * TN_BM);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per_split = ((m + splits - 1) / splits + TN_BM - 1) / TN_BM * TN_BM;
  splits = (m + m_per_split - 1) / m_per_split;
  dim3 grid((k + TN_BK - 1) / TN_BK, (n + TN_BN - 1) / TN_BN, splits);
  gemm_tn_kernel<<<grid, TN_T, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, c, ldc, m, n, k, m_per_split, colsum);
  GEOLDM_CHECK_LAUNCH("gemm_tn_kernel");
  return 0;
}
extern "C" int geoldm_gemm_tn(const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n, int k,
                              void* stream) {
  return geoldm_gemm_tn_bias(a, lda, b, ldb, c, ldc, nullptr, m, n, k, stream);
}

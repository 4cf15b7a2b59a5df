// Fused per-tile edge kernel, fp32 CUDA-core (FFMA) variant  —  GEOLDM_MMA_FP32_SIMT.
//
// Replaces, for one EGNN block, the reference's
//   coord2diff (egnn/egnn_new.py:249-255)  +  GCL.edge_model (:30-45)  +  unsorted_segment_sum (:258-274)
// or
//   coord2diff + EquivariantUpdate.coord_model (:86-99) + unsorted_segment_sum
// without ever materialising an [E, nf] edge tensor in HBM.
//
// One CTA = one tile of <= 64 consecutive edge rows (sorted by receiver i, sender j != i).
//   prologue : per-row geometry (r_ij, d0_ij, u_ij) into registers / shared memory
//   main loop: K-slabs of 32: A[row][k] = SiLU(P_i[k] + Q_j[k] + w_r[k] r + w_d[k] d0) generated into
//              shared memory (k-major), W2^T slab streamed with cp.async (double buffered),
//              64 x H x 32 register-tiled FFMA (8 x (H/32) accumulators per thread)
//   epilogue : m = SiLU(acc + b2); row-dot with w_att / w6 (shuffle + smem reduce);
//              GCL : e = m * sigmoid(dot + b_att) -> smem tile -> per-column sequential segment sum
//                    over j -> one atomicAdd per (receiver, column) and tile
//              EQV : delta = u * tanh(dot) * coords_range -> segment sum -> atomicAdd(xagg)
// A receiver's rows span at most two tiles when n-1 <= 64, so the two partial sums commute and the
// result is run-to-run deterministic for QM9-sized molecules.
#include "common.cuh"

namespace geoldm {

namespace {
constexpr int TM = 64;   // edge rows per tile
constexpr int BK = 32;   // k-slab
constexpr int NT = 256;  // threads per CTA

template <int H>
struct ColMap {
  static constexpr int WCOLS = H / 4;  // columns owned by one warp-column (4 warp-columns)
  static constexpr int VEC = (WCOLS % 32 == 0) ? 4 : ((WCOLS % 16 == 0) ? 2 : 1);
  static constexpr int Q = WCOLS / (8 * VEC);
  static constexpr int CN = VEC * Q;  // accumulator columns per thread
  __device__ __forceinline__ static int col(int wc, int cg, int q, int v) {
    return wc * WCOLS + q * 8 * VEC + cg * VEC + v;
  }
};

template <int H>
constexpr size_t edge_simt_smem_bytes() {
  size_t pipe = (size_t)2 * BK * (TM + H) * sizeof(float);
  size_t etile = (size_t)TM * H * sizeof(float);
  size_t main_bytes = pipe > etile ? pipe : etile;
  // + s_i[TM] + s_u[TM*3] + red[4*TM] + gs[TM]
  return main_bytes + (size_t)(TM + 3 * TM + 4 * TM + TM) * sizeof(float);
}

struct EdgeSimtArgs {
  const float* pq;     // [N][2H]
  const float* x;      // [N][3] block-entry coordinates
  const float* x0;     // [N][3] EGNN-entry coordinates
  const int* edge_i;
  const int* edge_j;
  const int* tile_row;
  const float* w_rd;   // [2][H]
  const float* w2t;    // [H][H]
  const float* b2;     // [H]
  const float* w_out;  // [H]
  const float* b_out;  // [1] or null
  float* out;          // agg [N][H] or xagg [N][3]
  float norm_constant, coords_range;
  int attention, use_tanh;
};

template <int H, bool EQUIV>
__global__ void __launch_bounds__(NT, 2) edge_simt_kernel(const EdgeSimtArgs a) {
  using CM = ColMap<H>;
  constexpr int CN = CM::CN, VEC = CM::VEC, Q = CM::Q;
  constexpr int NS = H / BK;

  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [2][BK][TM]
  float* Ws = smem + 2 * BK * TM;         // [2][BK][H]
  float* Es = smem;                       // [TM][H]  (aliases the pipeline buffers after the main loop)
  constexpr size_t MAIN = (2 * BK * (TM + H) > TM * H) ? (size_t)2 * BK * (TM + H) : (size_t)TM * H;
  int* s_i = reinterpret_cast<int*>(smem + MAIN);  // [TM]
  float* s_u = smem + MAIN + TM;          // [TM][3]
  float* red = s_u + 3 * TM;              // [4][TM]
  float* gs = red + 4 * TM;               // [TM]

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int wr = warp >> 2, wc = warp & 3;  // warp grid 2 (rows) x 4 (cols)
  const int rg = lane >> 3, cg = lane & 7;  // lane grid 4 (rows) x 8 (cols)

  const int row0 = a.tile_row[blockIdx.x];
  const int nrows = a.tile_row[blockIdx.x + 1] - row0;

  // ---- per-row metadata: generator thread (grow, kc) owns row `grow` for the whole tile --------
  const int grow = t & (TM - 1);
  const int kc = t >> 6;  // 0..3 : which 8-wide k chunk of each slab this thread generates
  const bool gvalid = grow < nrows;
  int gi = 0, gj = 0;
  float gr = 0.f, gd0 = 0.f;
  if (gvalid) {
    gi = a.edge_i[row0 + grow];
    gj = a.edge_j[row0 + grow];
    EdgeGeom g = edge_geom(a.x, a.x0, gi, gj, a.norm_constant);
    gr = g.r;
    gd0 = g.d0;
    if (kc == 0) {
      s_i[grow] = gi;
      if (EQUIV) {
        s_u[3 * grow] = g.ux;
        s_u[3 * grow + 1] = g.uy;
        s_u[3 * grow + 2] = g.uz;
      }
    }
  } else if (kc == 0) {
    s_i[grow] = -1;
  }
  const float* Pi = a.pq + (size_t)gi * (2 * H) + kc * 8;
  const float* Qj = a.pq + (size_t)gj * (2 * H) + H + kc * 8;

  float acc[8][CN];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < CN; ++c) acc[r][c] = 0.f;

  auto load_w = [&](int s, int buf) {
    const float* src = a.w2t + (size_t)s * BK * H;
    float* dst = Ws + buf * BK * H;
    for (int c = t; c < BK * H / 4; c += NT) cp_async16(dst + 4 * c, src + 4 * c);
    cp_async_commit();
  };
  float4 pf[4];  // prefetched P (2) and Q (2) for the next slab
  auto gen_load = [&](int s) {
    if (gvalid) {
      const float4* p = reinterpret_cast<const float4*>(Pi + s * BK);
      const float4* q = reinterpret_cast<const float4*>(Qj + s * BK);
      pf[0] = __ldg(p);
      pf[1] = __ldg(p + 1);
      pf[2] = __ldg(q);
      pf[3] = __ldg(q + 1);
    }
  };
  auto gen_store = [&](int s, int buf) {
    float* dst = As + buf * BK * TM + (kc * 8) * TM + grow;
    if (gvalid) {
      const float4* wr4 = reinterpret_cast<const float4*>(a.w_rd + s * BK + kc * 8);
      const float4* wd4 = reinterpret_cast<const float4*>(a.w_rd + H + s * BK + kc * 8);
      float p[8] = {pf[0].x, pf[0].y, pf[0].z, pf[0].w, pf[1].x, pf[1].y, pf[1].z, pf[1].w};
      float q[8] = {pf[2].x, pf[2].y, pf[2].z, pf[2].w, pf[3].x, pf[3].y, pf[3].z, pf[3].w};
      float4 w0 = __ldg(wr4), w1 = __ldg(wr4 + 1), d0 = __ldg(wd4), d1 = __ldg(wd4 + 1);
      float wrv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float wdv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = p[e] + q[e];
        v = fmaf(wrv[e], gr, v);
        v = fmaf(wdv[e], gd0, v);
        dst[e * TM] = silu(v);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[e * TM] = 0.f;
    }
  };

  load_w(0, 0);
  gen_load(0);
  gen_store(0, 0);

  for (int s = 0; s < NS; ++s) {
    const int buf = s & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (s + 1 < NS) {
      load_w(s + 1, buf ^ 1);
      gen_load(s + 1);
    }
    const float* Ab = As + buf * BK * TM + wr * 32 + rg * 4;
    const float* Wb = Ws + buf * BK * H;
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(Ab + k * TM);
      float4 a1 = *reinterpret_cast<const float4*>(Ab + k * TM + 16);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[CN];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float* bp = Wb + k * H + CM::col(wc, cg, q, 0);
        if (VEC == 4) {
          float4 b4 = *reinterpret_cast<const float4*>(bp);
          bv[q * VEC + 0] = b4.x; bv[q * VEC + 1] = b4.y; bv[q * VEC + 2] = b4.z; bv[q * VEC + 3] = b4.w;
        } else if (VEC == 2) {
          float2 b2 = *reinterpret_cast<const float2*>(bp);
          bv[q * VEC + 0] = b2.x; bv[q * VEC + 1] = b2.y;
        } else {
          bv[q * VEC] = *bp;
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < CN; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    if (s + 1 < NS) gen_store(s + 1, buf ^ 1);
  }

  // ---- epilogue -------------------------------------------------------------------------------
  // thread rows: wr*32 + rq*16 + rg*4 + v  (r = rq*4 + v);  thread cols: CM::col(wc, cg, q, v)
  float dot[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) dot[r] = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c = CM::col(wc, cg, q, v);
      const float bias = __ldg(a.b2 + c);
      const float wo = __ldg(a.w_out + c);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float m = silu(acc[r][q * VEC + v] + bias);
        acc[r][q * VEC + v] = m;
        dot[r] = fmaf(wo, m, dot[r]);
      }
    }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float d = dot[r];
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    dot[r] = d;
  }
  if (cg == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = wr * 32 + (r >> 2) * 16 + rg * 4 + (r & 3);
      red[wc * TM + row] = dot[r];
    }
  }
  __syncthreads();  // also: every warp is done reading As/Ws -> Es may alias them
  if (t < TM) {
    float d = (red[t] + red[TM + t]) + (red[2 * TM + t] + red[3 * TM + t]);
    float g;
    if (EQUIV) {
      g = a.use_tanh ? tanhf(d) : d;
    } else {
      g = a.attention ? sigmoidf_(d + __ldg(a.b_out)) : 1.0f;
    }
    gs[t] = (t < nrows) ? g : 0.f;
  }
  __syncthreads();

  if (!EQUIV) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = wr * 32 + (r >> 2) * 16 + rg * 4 + (r & 3);
      const float g = gs[row];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float* ep = Es + row * H + CM::col(wc, cg, q, 0);
        if (VEC == 4) {
          *reinterpret_cast<float4*>(ep) = make_float4(acc[r][q * 4] * g, acc[r][q * 4 + 1] * g,
                                                       acc[r][q * 4 + 2] * g, acc[r][q * 4 + 3] * g);
        } else if (VEC == 2) {
          *reinterpret_cast<float2*>(ep) = make_float2(acc[r][q * 2] * g, acc[r][q * 2 + 1] * g);
        } else {
          *ep = acc[r][q] * g;
        }
      }
    }
    __syncthreads();
    if (t < H) {
      float run = 0.f;
      int cur = s_i[0];
      for (int r = 0; r < nrows; ++r) {
        const int i = s_i[r];
        if (i != cur) {
          atomicAdd(a.out + (size_t)cur * H + t, run);
          run = 0.f;
          cur = i;
        }
        run += Es[r * H + t];
      }
      if (nrows > 0) atomicAdd(a.out + (size_t)cur * H + t, run);
    }
  } else {
    if (t < 3) {
      float run = 0.f;
      int cur = s_i[0];
      for (int r = 0; r < nrows; ++r) {
        const int i = s_i[r];
        if (i != cur) {
          atomicAdd(a.out + (size_t)cur * 3 + t, run);
          run = 0.f;
          cur = i;
        }
        float v = __fmul_rn(s_u[3 * r + t], gs[r]);   // (u * tanh(s)) * range, reference association
        if (a.use_tanh) v = __fmul_rn(v, a.coords_range);
        run += v;
      }
      if (nrows > 0) atomicAdd(a.out + (size_t)cur * 3 + t, run);
    }
  }
}

template <int H>
int launch_h(const EdgeSimtArgs& args, bool equiv, int n_tile, cudaStream_t st) {
  constexpr size_t smem = edge_simt_smem_bytes<H>();
  static DeviceOnce once;
  bool fresh;
  const int slot = device_slot(once, fresh);
  if (fresh) {
    cudaFuncSetAttribute(edge_simt_kernel<H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(edge_simt_kernel<H, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    once.done[slot] = true;
  }
  if (equiv)
    edge_simt_kernel<H, true><<<n_tile, NT, smem, st>>>(args);
  else
    edge_simt_kernel<H, false><<<n_tile, NT, smem, st>>>(args);
  GEOLDM_CHECK_LAUNCH("edge_simt_kernel");
  return 0;
}
}  // namespace

int launch_edge_simt(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                     const float* pq, const float* x, const float* x0, float* out, cudaStream_t st) {
  GEOLDM_REQUIRE(b.tile_m == TM, "edge_simt: batch tile_m=%d, kernel needs %d", b.tile_m, TM);
  if (b.n_tile == 0) return 0;
  EdgeSimtArgs a;
  a.pq = pq; a.x = x; a.x0 = x0;
  a.edge_i = b.edge_i; a.edge_j = b.edge_j; a.tile_row = b.tile_row;
  a.w_rd = w.w_rd; a.w2t = w.w2t; a.b2 = w.b2; a.w_out = w.w_out; a.b_out = w.b_out;
  a.out = out;
  a.norm_constant = cfg.norm_constant; a.coords_range = cfg.coords_range;
  a.attention = cfg.attention; a.use_tanh = cfg.tanh;
  GEOLDM_REQUIRE(equiv || !cfg.attention || w.b_out != nullptr, "edge_simt: attention needs b_out");
  switch (cfg.hidden_nf) {
    case 32: return launch_h<32>(a, equiv, b.n_tile, st);
    case 64: return launch_h<64>(a, equiv, b.n_tile, st);
    case 128: return launch_h<128>(a, equiv, b.n_tile, st);
    case 192: return launch_h<192>(a, equiv, b.n_tile, st);
    case 256: return launch_h<256>(a, equiv, b.n_tile, st);
    default: set_error("edge_simt: unsupported hidden_nf %d (32/64/128/192/256)", cfg.hidden_nf); return -1;
  }
}

}  // namespace geoldm

// tcgen05 / TMEM tensor-core kernels (GEOLDM_MMA_3XTF32, GEOLDM_MMA_TF32).
//
// One persistent, warp-specialised kernel template computes  D[128 x H] = A[128 x K] * W[H x K]^T  per tile on
// the 5th-generation tensor cores with fp32-equivalent products (3xTF32: a_hi*w_hi + a_hi*w_lo + a_lo*w_hi,
// fp32 accumulation in TMEM) and differs only in how A is produced and how D is consumed:
//
//   MODE_GCL   A[row][k] = SiLU(P_i[k] + Q_j[k] + w_r[k] r_ij + w_d[k] d0_ij)   (GCL.edge_model, egnn_new.py:30-45)
//              epilogue: m = SiLU(D + b2); g = sigmoid(w_att.m + b_att); agg_i += sum_j m*g   (:38-44, :258-267)
//   MODE_EQUIV same A with the coord_mlp weights                                 (EquivariantUpdate, :86-99)
//              epilogue: s = w6 . SiLU(D + b2); xagg_i += sum_j u_ij * tanh(s) * coords_range
//   MODE_DENSE A[row][k] = [a1 | a2/a2_div][row][k]  (node-level Linear: P|Q projections, node_mlp; :14-23,:53-55)
//              epilogue: out = epi(D + bias) (+ residual)
//   MODE_RAW   A[row][k] = P_{edge_i[row]}[k], epilogue dumps D  (self-test of descriptors / swizzle / pipeline)
//
// Roles (576 threads, 1 CTA per SM, ~200 KB shared memory, all 512 TMEM columns):
//   warps 0-7   epilogue  : tcgen05.ld of the fp32 accumulator; thread = (TMEM lane = edge row, column half);
//                           fused tail (bias, SiLU, head dot, gate, transposing segment sum, atomics)
//   warps 8-15  producers : generate the A k-slab (32 columns) straight into the canonical SWIZZLE_128B K-major
//                           shared-memory layout as tf32 hi / lo images; fence.proxy.async; mbarrier arrive
//   warp  16    loader    : cp.async.bulk (TMA engine) of the pre-swizzled hi|lo W stage, mbarrier complete_tx
//   warp  17    MMA       : one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=H/2, K=8) x 4 k-steps x 3
//                           terms per stage, tcgen05.commit releases the shared-memory stages / publishes D
// Pipelines: W stages = (k-slab, N-half) x 4, multicast across a cluster of CS CTAs (each CTA fetches 1/CS of a
// stage and the TMA engine writes it into every CTA of the cluster: L2->SM weight traffic / CS); A slabs 2 stages.
// TMEM: the K reduction of a tile is split in two halves accumulated in two 256-column regions (half the magnitude
// and half the number of round-toward-zero accumulations each); the epilogue first folds region 0 into region 1
// with a rounded fp32 add and releases region 0 at once, so the next tile's first K-half overlaps the fused tail of
// the current one.  No N^2 x nf tensor and no pre-activation ever reaches HBM.
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

// cycle-counter instrumentation of the MMA / producer threads: compiled in only with -DGEOLDM_TC_PROFILE
// (GEOLDM_TC_PROFILE=1 python -m geoldm_b200.build --force; read with scripts/tc_stats.py, GEOLDM_TC_DEBUG=32)
#ifdef GEOLDM_TC_PROFILE
#define TC_PROF(...) __VA_ARGS__
#else
#define TC_PROF(...)
#endif

namespace geoldm {
// cycle counters of the MMA-issuing thread (debug bit 32): where does the tensor pipe's feeder wait?
__device__ unsigned long long g_tc_stats[8];
__device__ unsigned long long g_tc_prod[2];
namespace {
using namespace tc;

constexpr int TM = 128;         // rows per tile (TMEM lanes)
constexpr int BK = 32;          // k-slab: 32 fp32/tf32 = 128 bytes per row = one swizzle row
constexpr int NTHREADS = 576;    // 8 epilogue + 8 producer warps + loader + MMA
constexpr int NWS = 3;          // W pipeline stages
constexpr int EPI_T = 256, PROD_T = 256;
constexpr int WARP_LOAD = 16, WARP_MMA = 17;
constexpr int MODE_GCL = 0, MODE_EQUIV = 1, MODE_DENSE = 2, MODE_RAW = 3;

struct TcArgs {
  // tiles
  int n_tile;               // row tiles
  int n_rows;               // total rows (edges, or nodes for dense)
  const int* tile_row;      // [n_tile+1] (edge modes) or nullptr (dense: tile t covers rows [128t, 128t+128))
  int n_blocks;             // column blocks of H outputs (dense: N_out / H); 1 otherwise
  int n_slabs;              // K / 32
  int terms;                // 3 (3xTF32) or 1 (single TF32 pass)
  // A operand sources
  const float* pq;          // [N][pq_ld] P in columns [0,H), Q in [H,2H)   (edge modes, RAW)
  int pq_ld;
  const float* x; const float* x0;
  const int* edge_i; const int* edge_j;
  const float* w_rd;        // [2][H]
  const float* a1; const float* a2; int k1, k2; float a2_div;   // dense
  // B operand: packed slabs  [block][slab][hi: H x 128 B | lo: H x 128 B], rows in SWIZZLE_128B order
  const float* w_pack;
  // epilogue
  const float* b2;          // [H] bias of the contraction (edge: second-layer bias; dense: bias [n_blocks*H] or null)
  const float* w_out;       // [H]
  const float* b_out;       // [1] or null
  const float* res;         // dense residual [M][ldo] or null
  float* out;               // agg [N][H] | xagg [N][3] | dense out [M][ldo] | raw dump [rows][H]
  int ldo; int epi;         // dense: row stride of out/res; 0 none, 1 SiLU, 2 residual
  float norm_constant, coords_range;
  int attention, use_tanh;
  int debug;                // timing experiments only (GEOLDM_TC_DEBUG): 1 skip W copies, 2 skip A generation, 4 skip tail
  float acc_scale;          // 1 + RZ_BIAS_PER_MMA * (#MMAs accumulated per output): see below
};

// The tensor core adds each MMA's partial sum into the fp32 TMEM accumulator with round-toward-zero (as on earlier
// generations: Fasi et al. 2021; Ootomo & Yokota 2022), so an output that went through n accumulating MMAs is
// shrunk towards zero by a systematic relative amount that we measured on B200 (scripts/tc_bias_probe.py) as
// 1.60e-8 * n, independent of H and of the operand distribution (H=64..256: 1.57, 1.59, 1.60, 1.60 e-8 per MMA;
// spread 0.42 of the mean).  The epilogue multiplies the accumulator by (1 + 1.60e-8 n) — fused into the bias
// FFMA — which removes the bias and leaves only the zero-mean part (6e-7 relative at n = 96).
constexpr float RZ_BIAS_PER_MMA = 1.60e-8f;


template <int H, int MODE>
struct Smem {
  static constexpr int NAS = (MODE == 0) ? 2 : 3;              // A pipeline stages (GCL spends 32 KB on transposition tiles)
  static constexpr uint32_t T_BYTES = (MODE == 0) ? 8u * 32 * 36 * 4 : 0u;   // 8 warps x [32 rows][36 floats]
  static constexpr uint32_t NH = H / 2;                       // W rows (= output columns) per stage
  static constexpr uint32_t W_IMG = NH * 128u;                // one tf32 image of a stage
  static constexpr uint32_t W_STAGE = 2u * W_IMG;             // hi + lo
  static constexpr uint32_t A_STAGE = 2u * TM * 128u;         // hi + lo
  static constexpr uint32_t OFF_W = 0;
  static constexpr uint32_t OFF_A = OFF_W + NWS * W_STAGE;
  static constexpr uint32_t OFF_T = OFF_A + NAS * A_STAGE;    // 8 warp-private [32][32] fp32 transposition tiles
  static constexpr uint32_t OFF_SI = OFF_T + T_BYTES;         // int   [128] receiver per row
  static constexpr uint32_t OFF_PS = OFF_SI + TM * 4;         // int   [129] piece starts
  static constexpr uint32_t OFF_DX = OFF_PS + 8 * 34 * 4;     // float [128][4] equiv deltas   (PS: 8 warps x 34 ints)
  static constexpr uint32_t OFF_DOT = OFF_DX + TM * 16;       // float [2][128] row-dot partials of the two column halves
  static constexpr uint32_t OFF_VEC = OFF_DOT + 2 * TM * 4;   // float [2][H] staged bias / head vectors
  static constexpr uint32_t OFF_CNT = OFF_VEC + 2 * H * 4;    // int   [8]
  static constexpr uint32_t OFF_BAR = OFF_CNT + 32;           // 3*NWS + 2*NAS + 4 mbarriers
  static constexpr uint32_t OFF_TMEM = OFF_BAR + (3 * NWS + 2 * NAS + 4) * 8;
  static constexpr uint32_t BYTES = OFF_TMEM + 16;
  static constexpr uint32_t ALLOC = BYTES + 1024;             // slack for manual 1024-byte alignment
};

template <int H, int MODE, int CS, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1) tc_kernel(const TcArgs a) {
  static_assert(!PAIR || CS == 2, "cta_group::2 needs a cluster of exactly two CTAs");
  using S = Smem<H, MODE>;
  constexpr int NAS = S::NAS;
  constexpr int NH = H / 2;
  extern __shared__ uint8_t smem_raw[];
  // same offset in every CTA of the cluster (multicast writes and barrier arrives address peers by offset)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* w_full = bars;                    // [NWS]
  uint64_t* w_empty = bars + NWS;             // [NWS]
  uint64_t* a_full = bars + 2 * NWS;                  // [NAS]
  uint64_t* a_empty = bars + 2 * NWS + NAS;           // [NAS]
  uint64_t* acc_full = bars + 2 * NWS + 2 * NAS;      // [2]  K-half 0 / 1 of the current tile complete
  uint64_t* acc_empty = bars + 2 * NWS + 2 * NAS + 2; // [2]  TMEM region 0 / 1 drained by the epilogue
  uint64_t* w_peer = bars + 2 * NWS + 2 * NAS + 4;    // [NWS] PAIR: the peer CTA's half of the W stage has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t cmask = (uint16_t)((1u << CS) - 1u);
  // work items: (tile, column block); padded so that all CTAs of a cluster run the same number of iterations on
  // the same column block (they share the W stream); padding tiles have zero rows
  const int n_tile_pad = (a.n_tile + CS - 1) / CS * CS;
  // PAIR: a work item is a super-tile of 256 rows (one 128-row tile per CTA of the pair)
  const int work_per_block = PAIR ? n_tile_pad / 2 : n_tile_pad;
  const int n_workers = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int worker = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int total_work = work_per_block * a.n_blocks;
  const int n_iter = (total_work + n_workers - 1) / n_workers;
  const int half_slabs = a.n_slabs / 2;

  if (tid == 0) {
    for (int s = 0; s < NWS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], PAIR ? 1 : CS);
      mbar_init(&w_peer[s], 1);
    }
    for (int s = 0; s < NAS; ++s) {
      mbar_init(&a_full[s], (PAIR ? 2 : 1) * PROD_T / 32);     // one arrive per producer warp (of both CTAs)
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], (PAIR ? 2 : 1) * EPI_T / 32);   // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) { if (PAIR) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  if (MODE == MODE_GCL || MODE == MODE_EQUIV) {     // stage the per-column vectors of the fused tail
    float* vec = reinterpret_cast<float*>(smem + S::OFF_VEC);
    for (int c = tid; c < H; c += NTHREADS) { vec[c] = a.b2[c]; vec[H + c] = a.w_out[c]; }
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();             // peers' barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_of = [&](int iter, int& tile, int& nb, int& row0, int& nrows) {
    const int work = iter * n_workers + worker;
    nb = work / work_per_block;
    tile = PAIR ? 2 * (work % work_per_block) + (int)crank : work % work_per_block;
    if (nb >= a.n_blocks || tile >= a.n_tile) {   // padding
      nb = nb >= a.n_blocks ? a.n_blocks - 1 : nb;
      row0 = 0; nrows = 0;
      return;
    }
    row0 = a.tile_row ? a.tile_row[tile] : tile * TM;
    nrows = (a.tile_row ? a.tile_row[tile + 1] : min(a.n_rows, row0 + TM)) - row0;
  };

  if (warp == WARP_LOAD) {
    // =========================== W-stage loader (TMA engine, multicast) ===================================
    if (PAIR) {
      // each CTA of the pair fetches ITS N-half of every k-slab (the hardware reads B rows [0,H/2) from the leader's
      // shared memory and [H/2,H) from the peer's); a second lane of the peer forwards "landed" to the leader
      if (lane == 0) {
        uint32_t wit = 0;
        for (int iter = 0; iter < n_iter; ++iter) {
          int tile, nb, row0, nrows;
          tile_of(iter, tile, nb, row0, nrows);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_pack) + (size_t)nb * a.n_slabs * 2 * S::W_STAGE;
          for (int s = 0; s < a.n_slabs; ++s, ++wit) {
            const int st = wit % NWS;
            mbar_wait(&w_empty[st], ((wit / NWS) & 1) ^ 1);
            if (a.debug & 1) { mbar_arrive(&w_full[st]); continue; }
            mbar_arrive_expect_tx(&w_full[st], S::W_STAGE);
            bulk_g2s(smem + S::OFF_W + st * S::W_STAGE, src + (size_t)(2 * s + crank) * S::W_STAGE, S::W_STAGE,
                     &w_full[st]);
          }
        }
      } else if (lane == 1 && crank == 1) {
        uint32_t wit = 0;
        for (int iter = 0; iter < n_iter; ++iter)
          for (int s = 0; s < a.n_slabs; ++s, ++wit) {
            const int st = wit % NWS;
            mbar_wait(&w_full[st], (wit / NWS) & 1);
            mbar_arrive_remote(&w_peer[st], 0);
          }
      }
    } else if (lane == 0) {
      uint32_t wit = 0;
      for (int iter = 0; iter < n_iter; ++iter) {
        int tile, nb, row0, nrows;
        tile_of(iter, tile, nb, row0, nrows);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_pack) + (size_t)nb * a.n_slabs * 2 * S::W_STAGE;
        for (int s = 0; s < 2 * a.n_slabs; ++s, ++wit) {       // (slab, N-half) stages
          const int st = wit % NWS;
          mbar_wait(&w_empty[st], ((wit / NWS) & 1) ^ 1);       // every CTA of the cluster released the stage
          if (a.debug & 1) { mbar_arrive(&w_full[st]); continue; }
          mbar_arrive_expect_tx(&w_full[st], S::W_STAGE);
          constexpr uint32_t PART = S::W_STAGE / CS;
          uint8_t* dst = smem + S::OFF_W + st * S::W_STAGE + crank * PART;
          const uint8_t* g = src + (size_t)s * S::W_STAGE + crank * PART;
          if (CS > 1) bulk_g2s_multicast(dst, g, PART, &w_full[st], cmask);
          else bulk_g2s(dst, g, PART, &w_full[st]);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================== MMA issuer ==============================================================
    if (PAIR) {
      // leader CTA only: M = 256 (128 rows from each CTA), N = H, operands and accumulators in both CTAs
      if (crank == 0) {
        const uint32_t idesc2 = make_idesc_tf32_m256(H);
        uint32_t wit = 0, ait = 0;
        TC_PROF(long long t_acc = 0; long long t_a = 0; long long t_w = 0; const long long t_begin = clock64();)
        for (int iter = 0; iter < n_iter; ++iter) {
          for (int kh = 0; kh < 2; ++kh) {
            TC_PROF(long long t0 = clock64();)
            mbar_wait_cluster(&acc_empty[kh], (iter & 1) ^ 1);
            TC_PROF(t_acc += clock64() - t0;)
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + kh * 256;
            for (int s = 0; s < half_slabs; ++s, ++ait, ++wit) {
              const int ast = ait % NAS, wst = wit % NWS;
              TC_PROF(t0 = clock64();)
              mbar_wait_cluster(&a_full[ast], (ait / NAS) & 1);
              TC_PROF(long long t1 = clock64();)
              mbar_wait(&w_full[wst], (wit / NWS) & 1);
              mbar_wait_cluster(&w_peer[wst], (wit / NWS) & 1);
              TC_PROF(long long t2 = clock64(); t_a += t1 - t0; t_w += t2 - t1;)
              tc_fence_after();
              if (lane == 0) {
                const uint32_t a_hi = smem_u32(smem + S::OFF_A + ast * S::A_STAGE);
                const uint32_t a_lo = a_hi + TM * 128;
                const uint32_t w_hi = smem_u32(smem + S::OFF_W + wst * S::W_STAGE);
                const uint32_t w_lo = w_hi + S::W_IMG;
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                  const uint64_t da_hi = make_smem_desc_sw128(a_hi + kk * 32), da_lo = make_smem_desc_sw128(a_lo + kk * 32);
                  const uint64_t dw_hi = make_smem_desc_sw128(w_hi + kk * 32), dw_lo = make_smem_desc_sw128(w_lo + kk * 32);
                  if (a.terms == 3) {
                    mma_tf32_pair(d_tmem, da_lo, dw_hi, idesc2, (s | kk) != 0);
                    mma_tf32_pair(d_tmem, da_hi, dw_lo, idesc2, 1);
                    mma_tf32_pair(d_tmem, da_hi, dw_hi, idesc2, 1);
                  } else {
                    mma_tf32_pair(d_tmem, da_hi, dw_hi, idesc2, (s | kk) != 0);
                  }
                }
                mma_commit_pair(&w_empty[wst], cmask);
                mma_commit_pair(&a_empty[ast], cmask);
                if (s == half_slabs - 1) mma_commit_pair(&acc_full[kh], cmask);
              }
              __syncwarp();
            }
          }
        }
        TC_PROF(if ((a.debug & 32) && lane == 0 && blockIdx.x == 0) {
          g_tc_stats[0] += (unsigned long long)(clock64() - t_begin);
          g_tc_stats[1] += (unsigned long long)t_acc;
          g_tc_stats[2] += (unsigned long long)t_a;
          g_tc_stats[3] += (unsigned long long)t_w;
          g_tc_stats[4] += 1ull;
          g_tc_stats[5] += (unsigned long long)n_iter;
        })
      }
    } else {
    const uint32_t idesc = (a.debug & 8) ? make_idesc_tf32(H) : make_idesc_tf32(NH);
    uint32_t wit = 0, ait = 0;
    for (int iter = 0; iter < n_iter; ++iter) {
      for (int kh = 0; kh < 2; ++kh) {
        mbar_wait(&acc_empty[kh], (iter & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kh * 256;
        for (int s = 0; s < half_slabs; ++s, ++ait) {
          const int ast = ait % NAS;
          mbar_wait(&a_full[ast], (ait / NAS) & 1);
          for (int nh = 0; nh < 2; ++nh, ++wit) {
            const int wst = wit % NWS;
            mbar_wait(&w_full[wst], (wit / NWS) & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_hi = smem_u32(smem + S::OFF_A + ast * S::A_STAGE);
              const uint32_t a_lo = a_hi + TM * 128;
              const uint32_t w_hi = smem_u32(smem + S::OFF_W + wst * S::W_STAGE);
              const uint32_t w_lo = w_hi + S::W_IMG;
              const uint32_t d = d_tmem + nh * NH;
#pragma unroll
              for (int kk = 0; kk < BK / 8; ++kk) {
                if ((a.debug & 8) && nh == 1) break;     // timing experiment: one N=H instruction instead of two N=H/2
                if ((a.debug & 16) && kk > 0) break;      // timing experiment: quarter of the MMAs
                const uint64_t da_hi = make_smem_desc_sw128(a_hi + kk * 32), da_lo = make_smem_desc_sw128(a_lo + kk * 32);
                const uint64_t dw_hi = make_smem_desc_sw128(w_hi + kk * 32), dw_lo = make_smem_desc_sw128(w_lo + kk * 32);
                if (a.terms == 3) {
                  mma_tf32(d, da_lo, dw_hi, idesc, (s | kk) != 0);
                  mma_tf32(d, da_hi, dw_lo, idesc, 1);
                  mma_tf32(d, da_hi, dw_hi, idesc, 1);
                } else {
                  mma_tf32(d, da_hi, dw_hi, idesc, (s | kk) != 0);
                }
              }
              if (CS > 1) mma_commit_multicast(&w_empty[wst], cmask);
              else mma_commit(&w_empty[wst]);
              if (nh == 1) {
                mma_commit(&a_empty[ast]);
                if (s == half_slabs - 1) mma_commit(&acc_full[kh]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
    }
  } else if (warp >= 8) {
    // =========================== A producers (256 threads) ==============================================
    const int pt = tid - EPI_T;
    const int chunk = pt & 7;        // 16-byte chunk: k = 4*chunk .. 4*chunk+3 inside the slab
    const int rbase = pt >> 3;       // rows rbase + 32 p
    uint32_t it = 0;
    TC_PROF(long long tp_wait = 0; long long tp_comp = 0; long long tp_fence = 0; long long tp_meta = 0;)
    for (int iter = 0; iter < n_iter; ++iter) {
      TC_PROF(const long long tm0 = clock64();)
      int tile, nb, row0, nrows;
      tile_of(iter, tile, nb, row0, nrows);
      const float* pP[4];
      const float* pQ[4];
      float rr[4], dd[4];
      bool valid[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int r = rbase + 32 * p;
        valid[p] = r < nrows;
        pP[p] = nullptr; pQ[p] = nullptr; rr[p] = 0.f; dd[p] = 0.f;
        if (valid[p]) {
          if (MODE == MODE_DENSE) {
            pP[p] = a.a1 + (size_t)(row0 + r) * a.k1 + 4 * chunk;
            pQ[p] = a.a2 ? a.a2 + (size_t)(row0 + r) * a.k2 + 4 * chunk : nullptr;
          } else {
            const int i = a.edge_i[row0 + r];
            pP[p] = a.pq + (size_t)i * a.pq_ld + 4 * chunk;
            if (MODE != MODE_RAW) {
              const int j = a.edge_j[row0 + r];
              pQ[p] = a.pq + (size_t)j * a.pq_ld + H + 4 * chunk;
              // squared distances only (the unit vector is needed by the EQUIV epilogue, not here);
              // same association as edge_geom / torch.sum(d**2, 1)
              const float* xi = a.x + 3 * (size_t)i;
              const float* xj = a.x + 3 * (size_t)j;
              const float* yi = a.x0 + 3 * (size_t)i;
              const float* yj = a.x0 + 3 * (size_t)j;
              const float dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
              const float ex = yi[0] - yj[0], ey = yi[1] - yj[1], ez = yi[2] - yj[2];
              rr[p] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
              dd[p] = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
            }
          }
        }
      }
      TC_PROF(tp_meta += clock64() - tm0;)
      for (int s = 0; s < a.n_slabs; ++s, ++it) {
        const int st = it % NAS;
        const int k0 = s * BK;
        float4 v[4];
        float4 q[4];
        TC_PROF(const long long tp0 = clock64();)
        // issue the global loads first; they are in flight while we wait for the stage to drain
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
          q[p] = v[p];
          if (valid[p]) {
            if (MODE == MODE_DENSE) {
              if (k0 < a.k1) v[p] = __ldg(reinterpret_cast<const float4*>(pP[p] + k0));
              else q[p] = __ldg(reinterpret_cast<const float4*>(pQ[p] + (k0 - a.k1)));
            } else {
              v[p] = __ldg(reinterpret_cast<const float4*>(pP[p] + k0));
              if (MODE != MODE_RAW) q[p] = __ldg(reinterpret_cast<const float4*>(pQ[p] + k0));
            }
          }
        }
        float4 wr = make_float4(0.f, 0.f, 0.f, 0.f), wd = wr;
        if (MODE == MODE_GCL || MODE == MODE_EQUIV) {
          wr = __ldg(reinterpret_cast<const float4*>(a.w_rd + k0 + 4 * chunk));
          wd = __ldg(reinterpret_cast<const float4*>(a.w_rd + H + k0 + 4 * chunk));
          // pull the next k-slab of the same P / Q rows into L1 (a tile touches ~40 distinct node rows = ~5 KB per
          // slab): costs no registers, and next iteration's loads no longer pay the L2 round trip
          if (s + 1 < a.n_slabs) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
              if (valid[p]) { prefetch_l1(pP[p] + k0 + BK); prefetch_l1(pQ[p] + k0 + BK); }
          }
        }
        mbar_wait(&a_empty[st], ((it / NAS) & 1) ^ 1);
        TC_PROF(const long long tp1 = clock64();)
        if (a.debug & 2) {
          __syncwarp();
          if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(&a_full[st], 0); else mbar_arrive(&a_full[st]); }
          continue;
        }
        uint8_t* a_hi = smem + S::OFF_A + st * S::A_STAGE;
        uint8_t* a_lo = a_hi + TM * 128;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int r = rbase + 32 * p;
          float e[4];
          if (MODE == MODE_GCL || MODE == MODE_EQUIV) {
            e[0] = silu(fmaf(wd.x, dd[p], fmaf(wr.x, rr[p], v[p].x + q[p].x)));
            e[1] = silu(fmaf(wd.y, dd[p], fmaf(wr.y, rr[p], v[p].y + q[p].y)));
            e[2] = silu(fmaf(wd.z, dd[p], fmaf(wr.z, rr[p], v[p].z + q[p].z)));
            e[3] = silu(fmaf(wd.w, dd[p], fmaf(wr.w, rr[p], v[p].w + q[p].w)));
            if (!valid[p]) e[0] = e[1] = e[2] = e[3] = 0.f;
          } else if (MODE == MODE_DENSE) {
            if (k0 < a.k1) {
              e[0] = v[p].x; e[1] = v[p].y; e[2] = v[p].z; e[3] = v[p].w;
            } else if (a.a2_div != 1.0f) {
              e[0] = __fdiv_rn(q[p].x, a.a2_div); e[1] = __fdiv_rn(q[p].y, a.a2_div);
              e[2] = __fdiv_rn(q[p].z, a.a2_div); e[3] = __fdiv_rn(q[p].w, a.a2_div);
            } else {
              e[0] = q[p].x; e[1] = q[p].y; e[2] = q[p].z; e[3] = q[p].w;
            }
          } else {
            e[0] = v[p].x; e[1] = v[p].y; e[2] = v[p].z; e[3] = v[p].w;
          }
          float4 hi, lo;
          split_tf32(e[0], hi.x, lo.x);
          split_tf32(e[1], hi.y, lo.y);
          split_tf32(e[2], hi.z, lo.z);
          split_tf32(e[3], hi.w, lo.w);
          const uint32_t off = sw128_off(r, chunk);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
        TC_PROF(const long long tp2 = clock64();)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(&a_full[st], 0); else mbar_arrive(&a_full[st]); }
        TC_PROF(if ((a.debug & 32) && tid == EPI_T && blockIdx.x == 0) {
          const long long tp3 = clock64();
          tp_wait += tp1 - tp0; tp_comp += tp2 - tp1; tp_fence += tp3 - tp2;
        })
      }
    }
    TC_PROF(if ((a.debug & 32) && tid == EPI_T && blockIdx.x == 0) {
      atomicAdd(&g_tc_stats[6], (unsigned long long)tp_wait);
      atomicAdd(&g_tc_stats[7], (unsigned long long)tp_comp);
      g_tc_prod[0] += (unsigned long long)tp_fence;
      g_tc_prod[1] += (unsigned long long)tp_meta;
    })
  } else {
    // ============= epilogue (warps 0-7: thread = (TMEM lane = row, column half)) ==========================
    const int r = (warp & 3) * 32 + lane;   // row / TMEM lane; a warp may only touch lanes 32*(warp%4)..+31
    const int hf = warp >> 2;               // column half handled by this thread
    constexpr int HC = H / 2;               // columns per half
    constexpr int NCH = HC / 32;            // 32-column chunks per half
    const int et = tid;                     // 0..255
    float* T = reinterpret_cast<float*>(smem + S::OFF_T);
    int* s_i = reinterpret_cast<int*>(smem + S::OFF_SI);
    int* s_ps = reinterpret_cast<int*>(smem + S::OFF_PS);
    float* s_dx = reinterpret_cast<float*>(smem + S::OFF_DX);
    float* s_dot = reinterpret_cast<float*>(smem + S::OFF_DOT);
    const float* s_b2 = reinterpret_cast<const float*>(smem + S::OFF_VEC) + hf * HC;
    const float* s_wo = s_b2 + H;
    int* s_cnt = reinterpret_cast<int*>(smem + S::OFF_CNT);
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + hf * HC;
    auto release_acc = [&](int region) {     // one arrive per warp, on the MMA-issuing (leader) CTA's barrier
      __syncwarp();
      if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(&acc_empty[region], 0); else mbar_arrive(&acc_empty[region]); }
    };
    for (int iter = 0; iter < n_iter; ++iter) {
      int tile, nb, row0, nrows;
      tile_of(iter, tile, nb, row0, nrows);
      const bool valid = r < nrows;
      int my_i = -1;
      float ux = 0.f, uy = 0.f, uz = 0.f;
      if (MODE == MODE_GCL || MODE == MODE_EQUIV) {
        if (valid) {
          my_i = a.edge_i[row0 + r];
          if (MODE == MODE_EQUIV && hf == 0) {
            EdgeGeom g = edge_geom(a.x, a.x0, my_i, a.edge_j[row0 + r], a.norm_constant);
            ux = g.ux; uy = g.uy; uz = g.uz;
          }
        }
        if (hf == 0) s_i[r] = my_i;
      }
      const uint32_t ph = iter & 1;
      mbar_wait(&acc_full[0], ph);
      mbar_wait(&acc_full[1], ph);
      tc_fence_after();
      if (a.debug & 4) {
        tc_fence_before();
        release_acc(0);
        release_acc(1);
        continue;
      }
      // ---- fold K-half 0 into K-half 1 (rounded fp32 add, RZ-bias compensation), release region 0 -----------
      const uint32_t taddr = tlane + 256;
#pragma unroll 1
      for (int cc = 0; cc < NCH; ++cc) {
        uint32_t v0[32], v1[32];
        tmem_ld32(tlane + cc * 32, v0);
        tmem_ld32(taddr + cc * 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c)
          v1[c] = __float_as_uint((__uint_as_float(v0[c]) + __uint_as_float(v1[c])) * a.acc_scale);
        tmem_st32(taddr + cc * 32, v1);
      }
      tmem_st_wait();
      tc_fence_before();
      release_acc(0);

      if (MODE == MODE_DENSE || MODE == MODE_RAW) {
        float* orow = a.out + (size_t)(row0 + r) * a.ldo + nb * H + hf * HC;
        const float* rrow = (MODE == MODE_DENSE && a.epi == 2) ? a.res + (size_t)(row0 + r) * a.ldo + nb * H + hf * HC : nullptr;
        const float* bias = (MODE == MODE_DENSE && a.b2) ? a.b2 + nb * H + hf * HC : nullptr;
#pragma unroll 1
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t v[32];
          tmem_ld32(taddr + cc * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              float o[4] = {__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]), __uint_as_float(v[c4 * 4 + 2]),
                            __uint_as_float(v[c4 * 4 + 3])};
              if (bias) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + c4 * 4));
                o[0] += b4.x; o[1] += b4.y; o[2] += b4.z; o[3] += b4.w;
              }
              if (MODE == MODE_DENSE && a.epi == 1) { o[0] = silu(o[0]); o[1] = silu(o[1]); o[2] = silu(o[2]); o[3] = silu(o[3]); }
              if (rrow) {
                const float4 rs = __ldg(reinterpret_cast<const float4*>(rrow + cc * 32 + c4 * 4));
                o[0] += rs.x; o[1] += rs.y; o[2] += rs.z; o[3] += rs.w;
              }
              *reinterpret_cast<float4*>(orow + cc * 32 + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
        tc_fence_before();
        release_acc(1);
      } else {
        // ---- pass 1: m = SiLU(D + b2), partial row dot with w_att / w6 over this thread's column half ----------
        float dot = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t v[32];
          tmem_ld32(taddr + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_b2 + cc * 32 + c4 * 4);
            const float4 w4 = *reinterpret_cast<const float4*>(s_wo + cc * 32 + c4 * 4);
            const float m0 = silu(__uint_as_float(v[c4 * 4 + 0]) + b4.x);
            const float m1 = silu(__uint_as_float(v[c4 * 4 + 1]) + b4.y);
            const float m2 = silu(__uint_as_float(v[c4 * 4 + 2]) + b4.z);
            const float m3 = silu(__uint_as_float(v[c4 * 4 + 3]) + b4.w);
            dot = fmaf(w4.x, m0, dot); dot = fmaf(w4.y, m1, dot); dot = fmaf(w4.z, m2, dot); dot = fmaf(w4.w, m3, dot);
            v[c4 * 4 + 0] = __float_as_uint(m0); v[c4 * 4 + 1] = __float_as_uint(m1);
            v[c4 * 4 + 2] = __float_as_uint(m2); v[c4 * 4 + 3] = __float_as_uint(m3);
          }
          if (MODE == MODE_GCL) tmem_st32(taddr + cc * 32, v);
        }
        // the two warps that share this lane quarter (column halves) exchange their partial dots: 64-thread barrier
        s_dot[hf * TM + r] = dot;
        if (MODE == MODE_GCL) tmem_st_wait();
        named_bar_sync(2 + (warp & 3), 64);
        const float full_dot = s_dot[r] + s_dot[TM + r];
        // receiver runs ("pieces") inside this warp's 32 rows; a receiver's rows span at most two warps when n-1 <= 32,
        // so every (receiver, column) gets at most two partial atomics: order-independent, hence deterministic
        const int prev_i = __shfl_up_sync(0xffffffffu, my_i, 1);
        const bool head = valid && (lane == 0 || prev_i != my_i);
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        const int nval = __popc(__ballot_sync(0xffffffffu, valid));
        const int npiece = __popc(hm);
        if (MODE == MODE_EQUIV) {
          tc_fence_before();
          release_acc(1);                  // accumulator no longer needed
          if (hf == 0) {
            float phi = a.use_tanh ? tanhf(full_dot) : full_dot;
            float dx = __fmul_rn(ux, phi), dy = __fmul_rn(uy, phi), dz = __fmul_rn(uz, phi);
            if (a.use_tanh) { dx = __fmul_rn(dx, a.coords_range); dy = __fmul_rn(dy, a.coords_range); dz = __fmul_rn(dz, a.coords_range); }
            s_dx[4 * r] = valid ? dx : 0.f; s_dx[4 * r + 1] = valid ? dy : 0.f; s_dx[4 * r + 2] = valid ? dz : 0.f;
            __syncwarp();
            // run heads sum their rows sequentially (ascending j, like the reference's scatter order)
            if (head) {
              const unsigned after = hm & ~((2u << lane) - 1u);           // heads strictly above this lane
              const int q1 = after ? (__ffs(after) - 1) : nval;
              float sx = 0.f, sy = 0.f, sz = 0.f;
              for (int q = lane; q < q1; ++q) {
                const float* d = s_dx + 4 * ((warp & 3) * 32 + q);
                sx += d[0]; sy += d[1]; sz += d[2];
              }
              atomicAdd(a.out + (size_t)my_i * 3, sx);
              atomicAdd(a.out + (size_t)my_i * 3 + 1, sy);
              atomicAdd(a.out + (size_t)my_i * 3 + 2, sz);
            }
            __syncwarp();                  // s_dx rows reused by the next tile
          }
          named_bar_sync(2 + (warp & 3), 64);   // s_dot reused by the next tile
        } else {
          float g = a.attention ? sigmoidf_(full_dot + __ldg(a.b_out)) : 1.0f;
          if (!valid) g = 0.f;
          // ---- pass 2: e = m * g; each warp transposes its own 32 rows x 16 columns through a private 2 KB smem
          //      slice (no CTA barriers) and its two half-warps take the column sums of alternate pieces ------------
          // row stride 36 floats: 16-byte aligned rows whose float4 stores (lane = row) and column reads (lane = column)
          // are both bank-conflict free without an XOR swizzle
          float* Tw = reinterpret_cast<float*>(smem + S::OFF_T) + warp * (32 * 36);
          int* psw = s_ps + warp * 34;        // this warp's piece starts (+ end sentinel)
          if (head) psw[__popc(hm & ((1u << lane) - 1u))] = lane;
          if (lane == 0) psw[npiece] = nval;
          __syncwarp();
#pragma unroll 1
          for (int cc = 0; cc < NCH; ++cc) {
            uint32_t v[32];
            tmem_ld32(taddr + cc * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              float4 e4 = make_float4(__uint_as_float(v[c4 * 4]) * g, __uint_as_float(v[c4 * 4 + 1]) * g,
                                      __uint_as_float(v[c4 * 4 + 2]) * g, __uint_as_float(v[c4 * 4 + 3]) * g);
              *reinterpret_cast<float4*>(Tw + lane * 36 + c4 * 4) = e4;
            }
            __syncwarp();
            for (int pc = 0; pc < npiece; ++pc) {        // lane = column; pieces in row order
              const int q0 = psw[pc], q1 = psw[pc + 1];
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
              const float* tp = Tw + q0 * 36 + lane;
              int q = q0;
              for (; q + 4 <= q1; q += 4, tp += 4 * 36) {
                s0 += tp[0];
                s1 += tp[36];
                s2 += tp[72];
                s3 += tp[108];
              }
              for (; q < q1; ++q, tp += 36) s0 += tp[0];
              const int pi = s_i[(warp & 3) * 32 + q0];   // written by this quarter's hf == 0 warp before the 64-thread barrier
              atomicAdd(a.out + (size_t)pi * H + hf * HC + cc * 32 + lane, (s0 + s1) + (s2 + s3));
            }
            __syncwarp();
          }
          tc_fence_before();
          release_acc(1);
          named_bar_sync(2 + (warp & 3), 64);   // s_dot reused by the next tile
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();             // no CTA exits while a peer may still multicast into it
  if (warp == WARP_MMA) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int H, int MODE, int CS, bool PAIR>
int launch_mode(const TcArgs& a, cudaStream_t st) {
  using S = Smem<H, MODE>;
  static DeviceOnce once;
  bool fresh;
  const int slot = device_slot(once, fresh);
  if (fresh) {
    cudaError_t e = cudaFuncSetAttribute(tc_kernel<H, MODE, CS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::ALLOC);
    if (e != cudaSuccess) {
      set_error("tc_kernel: cannot reserve %u bytes of shared memory: %s", S::ALLOC, cudaGetErrorString(e));
      return -2;
    }
    once.done[slot] = true;
  }
  const int sm_count = once.sm_count[slot];
  GEOLDM_REQUIRE(a.n_slabs % 2 == 0, "tc_kernel: K/32 = %d must be even", a.n_slabs);
  const int n_tile_pad = (a.n_tile + CS - 1) / CS * CS;
  const int work = n_tile_pad * a.n_blocks;     // in 128-row tiles (a PAIR work item covers two of them)
  if (a.n_tile == 0) return 0;
  int grid = work < sm_count ? work : sm_count;
  grid = grid / CS * CS;
  TcArgs args = a;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("GEOLDM_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  args.debug = dbg;
  args.acc_scale = a.terms == 3 ? 1.0f + RZ_BIAS_PER_MMA * (float)((a.n_slabs / 2) * (BK / 8) * 3) : 1.0f;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = S::ALLOC;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_kernel<H, MODE, CS, PAIR>, args);
  if (e != cudaSuccess) {
    set_error("tc_kernel launch: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int g_cluster = 2;   // W-multicast cluster size (1, 2 or 4); GEOLDM_TC_CLUSTER overrides
int g_pair = 1;      // 1: cta_group::2 MMAs over a CTA pair (default); GEOLDM_TC_PAIR=0 selects cta_group::1

template <int H, int MODE>
int launch_cs(const TcArgs& a, cudaStream_t st) {
  static bool read_env = false;
  if (!read_env) {
    const char* e = getenv("GEOLDM_TC_CLUSTER");
    if (e) g_cluster = atoi(e);
    e = getenv("GEOLDM_TC_PAIR");
    if (e) g_pair = atoi(e);
    read_env = true;
  }
  if (g_pair) return launch_mode<H, MODE, 2, true>(a, st);
  switch (g_cluster) {
    case 1: return launch_mode<H, MODE, 1, false>(a, st);
    case 4: return launch_mode<H, MODE, 4, false>(a, st);
    default: return launch_mode<H, MODE, 2, false>(a, st);
  }
}

template <int MODE>
int launch_h(int H, const TcArgs& a, cudaStream_t st) {
  switch (H) {
    case 64: return launch_cs<64, MODE>(a, st);
    case 128: return launch_cs<128, MODE>(a, st);
    case 192: return launch_cs<192, MODE>(a, st);
    case 256: return launch_cs<256, MODE>(a, st);
    default: set_error("tcgen05 kernels support hidden_nf 64/128/192/256, got %d", H); return -1;
  }
}
}  // namespace

int launch_edge_tc(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                   const float* pq, int pq_ld, const float* x, const float* x0, float* out, cudaStream_t st) {
  if (cfg.mma_mode == GEOLDM_MMA_3XF16) return launch_edge_tc16(cfg, w, b, equiv, pq, pq_ld, x, x0, nullptr, nullptr, nullptr, out, st);
  GEOLDM_REQUIRE(b.tile_m == TM, "edge_tc: batch tile_m=%d, kernel needs %d", b.tile_m, TM);
  GEOLDM_REQUIRE(w.tc_pack != nullptr, "edge_tc: tc_pack missing (weights not packed for the tensor-core path)");
  GEOLDM_REQUIRE(equiv || !cfg.attention || w.b_out != nullptr, "edge_tc: attention needs b_out");
  TcArgs a{};
  a.n_tile = b.n_tile; a.n_rows = b.n_edge; a.tile_row = b.tile_row; a.n_blocks = 1;
  a.n_slabs = cfg.hidden_nf / BK;
  a.terms = cfg.mma_mode == GEOLDM_MMA_TF32 ? 1 : 3;
  a.pq = pq; a.pq_ld = pq_ld; a.x = x; a.x0 = x0; a.edge_i = b.edge_i; a.edge_j = b.edge_j; a.w_rd = w.w_rd;
  a.w_pack = reinterpret_cast<const float*>(w.tc_pack);
  a.b2 = w.b2; a.w_out = w.w_out; a.b_out = w.b_out; a.out = out;
  a.norm_constant = cfg.norm_constant; a.coords_range = cfg.coords_range;
  a.attention = cfg.attention; a.use_tanh = cfg.tanh;
  return equiv ? launch_h<MODE_EQUIV>(cfg.hidden_nf, a, st) : launch_h<MODE_GCL>(cfg.hidden_nf, a, st);
}

// out[M][ldo] (column block nb: columns [nb*H, nb*H+H)) = epi(A * W_nb^T + bias) (+res);  A = [a1 | a2/a2_div]
int launch_linear_tc(int H, int terms, const float* a1, int k1, const float* a2, int k2, float a2_div,
                     const void* w_pack, int n_blocks, const float* bias, const float* res, int epi, float* out, int m,
                     cudaStream_t st, float* zero_buf) {
  if (terms == 16) return launch_linear_tc16(H, a1, k1, a2, k2, a2_div, w_pack, n_blocks, bias, res, epi, out, m, st, zero_buf);
  GEOLDM_REQUIRE(zero_buf == nullptr, "linear_tc: zero_buf is implemented by the fp16-split kernels only");
  GEOLDM_REQUIRE(k1 % BK == 0 && k2 % BK == 0 && k1 + k2 > 0, "linear_tc: k1=%d k2=%d must be multiples of %d", k1, k2, BK);
  GEOLDM_REQUIRE(w_pack != nullptr, "linear_tc: w_pack missing");
  TcArgs a{};
  a.n_tile = (m + TM - 1) / TM; a.n_rows = m; a.tile_row = nullptr; a.n_blocks = n_blocks;
  a.n_slabs = (k1 + k2) / BK; a.terms = terms;
  a.a1 = a1; a.a2 = a2; a.k1 = k1; a.k2 = k2; a.a2_div = a2_div;
  a.w_pack = reinterpret_cast<const float*>(w_pack);
  a.b2 = bias; a.res = res; a.epi = epi; a.out = out; a.ldo = n_blocks * H;
  return launch_h<MODE_DENSE>(H, a, st);
}

// self-test: out[rows][H] = pq[edge_i[row]][0:H] * W^T   (K = H)
int launch_tc_selftest(int H, int terms, const float* pq, const int* edge_i, const int* tile_row, int n_tile,
                       int n_rows, const void* w_pack, float* out, cudaStream_t st) {
  if (terms == 16) return launch_tc16_selftest(H, pq, edge_i, tile_row, n_tile, n_rows, w_pack, out, st);
  TcArgs a{};
  a.n_tile = n_tile; a.n_rows = n_rows; a.tile_row = tile_row; a.n_blocks = 1; a.n_slabs = H / BK; a.terms = terms;
  a.pq = pq; a.pq_ld = 2 * H; a.edge_i = edge_i; a.w_pack = reinterpret_cast<const float*>(w_pack); a.out = out; a.ldo = H;
  return launch_h<MODE_RAW>(H, a, st);
}

}  // namespace geoldm

extern "C" int geoldm_has_tcgen05(void) { return 1; }
// debug: read and reset the MMA-thread cycle counters {total, wait acc_empty, wait a_full, wait w_full, launches, tiles}
extern "C" int geoldm_tc_read_stats(unsigned long long* host_out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_out, geoldm::g_tc_stats, sizeof(unsigned long long) * 8);
  cudaMemcpyFromSymbol(host_out + 8, geoldm::g_tc_prod, sizeof(unsigned long long) * 2);
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(geoldm::g_tc_stats, zero, sizeof(zero));
  cudaMemcpyToSymbol(geoldm::g_tc_prod, zero, sizeof(unsigned long long) * 2);
  return 0;
}

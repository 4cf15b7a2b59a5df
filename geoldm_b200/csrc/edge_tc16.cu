// tcgen05 / TMEM tensor-core kernels, fp16-split arithmetic (GEOLDM_MMA_3XF16).
//
// Same fused kernels as edge_tc.cu (GCL / EQUIV / DENSE / RAW, same roles, same shared-memory images, CTA pairs with
// cta_group::2 MMAs), with the fp32-equivalent product formed from fp16 halves instead of tf32 halves:
//     a = a_hi + a_lo,  w * 2^e = w_hi + w_lo   (all four fp16, 11 significant bits each, like tf32)
//     a * w = 2^-e (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi)                 fp32 accumulation in TMEM
// kind::f16 MMAs run at twice the tf32 rate and a 128-byte swizzle row holds 64 K-columns instead of 32, so a tile
// needs half the k-slabs, half the MMA time and half the operand bytes per K.  That frees the second 256-column
// TMEM region (the tf32 kernel needs it for its K-split): here consecutive tiles ALTERNATE between the two regions,
// so the fused tail of tile t overlaps the MMAs of tile t+1 and the A generation of tile t+2.
// The weight images carry a power-of-two scale 2^e (chosen by geoldm_tc_pack16 so that max|w| 2^e is in [2^13, 2^14)):
// it keeps w_lo out of the fp16 subnormal range; the epilogue folds 2^-e into the round-toward-zero compensation
// factor.  Activations are not scaled: |a| < 65504 is required for full accuracy (larger values saturate, they do not become
// inf / NaN); a_lo below 2^-14 is subnormal: absolute error <= 2^-25.
#include <cuda.h>
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

#ifdef GEOLDM_TC_PROFILE
#define TC_PROF(...) __VA_ARGS__
#else
#define TC_PROF(...)
#endif

namespace geoldm {
// cycle counters of one MMA / producer / epilogue thread of CTA 0 (GEOLDM_TC_PROFILE builds only):
// [0] MMA total [1] wait acc_empty [2] wait a_full [3] wait w [4] launches [5] tiles
// [6] producer wait a_empty (+loads) [7] producer compute+store [8] producer fence+arrive [9] producer metadata
// [10] epilogue wait acc_full [11] pass 1 [12] barrier + gate [13] pass 2 [14] epilogue metadata
__device__ unsigned long long g_tc16_stats[16];
namespace {
using namespace tc;

constexpr int TM = 128;          // rows per tile (TMEM lanes)
constexpr int BK = 64;           // k-slab: 64 fp16 = 128 bytes per row = one swizzle row
// K order inside a slab: the 16-byte operand chunk c (8 fp16) holds the slab's columns 4c..4c+3 and 32+4c..32+4c+3, in the
// A images and in the packed W images alike (a common K permutation leaves the product unchanged).  A producer thread
// then reads two 16-byte pieces that are 128 bytes apart and the 8 chunk-lanes of a row cover one contiguous 128-byte
// line per load instruction (half the L1 wavefronts of the 32-bytes-per-lane order).
constexpr int KH = BK / 2;
#ifndef GEOLDM_EQUIV_SPS
#define GEOLDM_EQUIV_SPS 1
#endif
// Warp roles per mode.  The A generator is latency bound (two dependent global-load phases per k-slab and thread), so
// EQUIV, DENSE and RAW run 16 producer warps (2 rows x 8 columns per thread and slab, one load phase) next to the 8
// epilogue warps (832 threads, 72 registers); GCL keeps 8 + 8 (576 threads, 96 registers): its tail (second SiLU, gate,
// segment sum) is as long as its A generation and its 4-row producers spill at 72 registers (measured slower).
// Epilogue warp w owns TMEM lanes 32 (w % 4) .. +31 (= tile rows) and column half w / 4.
template <int MODE>
struct Roles {
#ifndef GEOLDM_GCL_PROD_W
#define GEOLDM_GCL_PROD_W 8
#endif
  static constexpr int EPI_W = 8;
#ifndef GEOLDM_EQUIV_PROD_W
#define GEOLDM_EQUIV_PROD_W 16
#endif
  static constexpr int PROD_W = (MODE == 0) ? GEOLDM_GCL_PROD_W : (MODE == 1) ? GEOLDM_EQUIV_PROD_W : 16;
  static constexpr int WARP_LOAD = EPI_W + PROD_W, WARP_MMA = WARP_LOAD + 1;
  static constexpr int NTHREADS = 32 * (EPI_W + PROD_W + 2);
  static constexpr int EPI_T = 32 * EPI_W;
  static constexpr int ROWS_PT = 128 / (4 * PROD_W);          // consecutive tile rows per producer thread: 4 or 2
  static constexpr int NHALF = EPI_W / 4;                     // column halves split over distinct epilogue warps
};
constexpr int MODE_GCL = 0, MODE_EQUIV = 1, MODE_DENSE = 2, MODE_RAW = 3;
// per-tile staging of the first-layer projections (TMA tensor copies): one box of STAGE_PBOX receiver rows and one of
// STAGE_QBOX sender rows x 64 fp32 columns per k-slab
constexpr uint32_t STAGE_PBOX = 16, STAGE_QBOX = 64;
constexpr uint32_t PACK_HDR = 128;   // bytes: float inv_scale at offset 0
constexpr float RZ_BIAS_PER_MMA = 1.60e-8f;   // see edge_tc.cu (round-toward-zero accumulation of the tensor core)

struct Args {
  int n_tile, n_rows;
  const int* tile_row;
  int n_blocks;
  int n_slabs;              // K / 64
  const float* pq; int pq_ld;
  const float* x; const float* x0;
  const int* edge_i; const int* edge_j;
  const float* w_rd;
  const float* r_edge; const float* d0_edge;   // [E] precomputed squared distances (or null: computed from x / x0)
  const float4* u_edge;                        // [E] precomputed (x_i - x_j) / (|x_i - x_j| + c) (or null)
  const int4* tile_meta;                       // [n_tile] {first receiver, first sender, staged, 0} (or null: gather path)
  const float* a1; const float* a2; int k1, k2; float a2_div;
  const uint8_t* w_pack;    // header + [block][slab][N-half][hi image | lo image]
  const float* b2; const float* w_out; const float* b_out; const float* res;
  float* out; int ldo; int epi;
  float* zero_buf;          // DENSE residual epilogue: same layout as out, cleared on the way (the consumed agg buffer)
  float norm_constant, coords_range;
  int attention, use_tanh;
  float rz_scale;           // 1 + RZ_BIAS_PER_MMA * (#MMAs accumulated per output)
  // RAW mode, optional: bit pattern of max|A| over the whole operand (device scalar).  The producers multiply A by the
  // power of two that puts that maximum in [2^13, 2^14) and the epilogue divides it out again (both exact): gradient
  // operands, which are 1 / batch-size small, stay inside the normal fp16 range of the hi | lo split.
  const unsigned* a_amax_bits;
};

// power-of-two exponent e with amax 2^e in [2^13, 2^14) (0 for amax == 0), from the fp32 bit pattern of amax
__device__ __forceinline__ int amax_exponent(unsigned bits) {
  const float amax = __uint_as_float(bits);
  int e = 0;
  if (amax > 0.f) { int ex; frexpf(amax, &ex); e = 14 - ex; }
  return max(-100, min(100, e));
}

template <int H, int MODE>
struct Smem {
  // A pipeline: NAS stages of SPS k-slabs each (one wait / proxy fence / arrive per STAGE), W pipeline: NWS k-slab stages.
  // EQUIV has no transposition tiles and can spend that shared memory on two-slab A stages (half the hand-offs per tile).
  static constexpr int SPS = (MODE == 1 && (H / 64) % 2 == 0) ? GEOLDM_EQUIV_SPS : 1;
  static constexpr int NAS = (MODE == 1) ? (SPS == 2 ? 2 : 3) : 2;
#ifndef GEOLDM_EDGE_NWS
#define GEOLDM_EDGE_NWS 2
#endif
  static constexpr int NWS = (SPS == 2) ? 2 : ((MODE <= 1) ? GEOLDM_EDGE_NWS : 3);
  // warp-private [32 rows][36 floats] transposition tiles: GCL 8 warps (segment sum), DENSE / RAW 4 warps (coalesced
  // output rows); EQUIV needs none
  static constexpr uint32_t T_BYTES = (MODE == 1) ? 0u : 8u * 32 * 36 * 4;
  static constexpr uint32_t NH = H / 2;
  static constexpr uint32_t W_IMG = NH * 128u;
  static constexpr uint32_t W_STAGE = 2u * W_IMG;
  static constexpr uint32_t A_SLAB = 2u * TM * 128u;          // hi | lo images of one 64-column k-slab
  static constexpr uint32_t A_STAGE = SPS * A_SLAB;
  static constexpr uint32_t OFF_W = 0;
  static constexpr uint32_t OFF_A = OFF_W + NWS * W_STAGE;
  // staged first-layer projections (edge modes): NPS stages of [STAGE_PBOX + STAGE_QBOX rows][64 fp32], filled by TMA
  static constexpr int NPS = (MODE <= 1) ? 2 : 0;
  static constexpr uint32_t PQ_STAGE = (STAGE_PBOX + STAGE_QBOX) * 256u;
  static constexpr uint32_t OFF_PQ = OFF_A + NAS * A_STAGE;
  static constexpr uint32_t OFF_T = OFF_PQ + NPS * PQ_STAGE;
  static constexpr uint32_t OFF_SI = OFF_T + T_BYTES;
  static constexpr uint32_t OFF_PS = OFF_SI + TM * 4;
  static constexpr uint32_t OFF_DX = OFF_PS + 8 * 34 * 4;
  static constexpr uint32_t OFF_DOT = OFF_DX + TM * 16;
  static constexpr uint32_t OFF_VEC = OFF_DOT + 2 * TM * 4;
  static constexpr uint32_t OFF_WRD = OFF_VEC + 2 * H * 4;    // float [2][H]: distance columns of the first edge layer
  static constexpr uint32_t OFF_BAR = OFF_WRD + 2 * H * 4;
  static constexpr uint32_t OFF_TMEM = OFF_BAR + (3 * NWS + 2 * NAS + 4 + 2 * NPS) * 8;
  static constexpr uint32_t BYTES = OFF_TMEM + 16;
  static constexpr uint32_t ALLOC = BYTES + 1024;
};

// kind::f16 (fp16 inputs), fp32 accumulate, A and B K-major, M = 256 over a CTA pair, N = n
__host__ __device__ constexpr uint32_t make_idesc_f16_m256(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// two fp32 -> packed fp16x2 (first argument in the low half), round to nearest, overflow saturates to +-65504 instead of
// becoming inf: an out-of-range activation then degrades the product instead of poisoning the tile with NaN
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo_half, float hi_half) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}
// 8 fp32 -> 8 fp16 hi + 8 fp16 lo (lo = fp16(v - hi), exact subtraction), packed for one 16-byte swizzle chunk
__device__ __forceinline__ void split_f16x8(const float (&e)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_f16x2_sat(e[2 * i], e[2 * i + 1]);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
    l[i] = pack_f16x2_sat(e[2 * i] - hf.x, e[2 * i + 1] - hf.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// packed pair -> fp16x2 hi and fp16x2 lo (lo = fp16(v - hi), exact subtraction)
__device__ __forceinline__ void split_f16x2(f32x2 e, uint32_t& hi, uint32_t& lo) {
  float e0, e1;
  upk2(e, e0, e1);
  hi = pack_f16x2_sat(e0, e1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(e), "l"(pk2(hf.x, hf.y)));
  upk2(d, e0, e1);
  lo = pack_f16x2_sat(e0, e1);
}

template <int H, int MODE>
__global__ void __launch_bounds__(Roles<MODE>::NTHREADS, 1) tc16_kernel(const Args a, const __grid_constant__ CUtensorMap tm_p,
                                                                        const __grid_constant__ CUtensorMap tm_q) {
  using S = Smem<H, MODE>;
  using R = Roles<MODE>;
  constexpr int NTHREADS = R::NTHREADS, EPI_T = R::EPI_T, WARP_LOAD = R::WARP_LOAD, WARP_MMA = R::WARP_MMA;
  constexpr int ROWS_PT = R::ROWS_PT, NHALF = R::NHALF;
  constexpr int NAS = S::NAS, NWS = S::NWS, SPS = S::SPS, NPS = S::NPS > 0 ? S::NPS : 1;   // (NPS: no zero divisor in dead code)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* w_full = bars;                            // [NWS]
  uint64_t* w_empty = bars + NWS;                     // [NWS]
  uint64_t* a_full = bars + 2 * NWS;                  // [NAS]
  uint64_t* a_empty = bars + 2 * NWS + NAS;           // [NAS]
  uint64_t* acc_full = bars + 2 * NWS + 2 * NAS;      // [2]  accumulator region r holds a complete tile
  uint64_t* acc_empty = bars + 2 * NWS + 2 * NAS + 2; // [2]  region r drained by the epilogue
  uint64_t* w_peer = bars + 2 * NWS + 2 * NAS + 4;    // [NWS] the peer CTA's N-half of the W stage has landed
  uint64_t* pq_full = bars + 3 * NWS + 2 * NAS + 4;   // [NPS] staged projection rows of a k-slab have landed (TMA)
  uint64_t* pq_empty = pq_full + NPS;                 // [NPS] ... have been read by every producer warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);

  TC_PROF(const long long t_entry = clock64();)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  constexpr uint16_t cmask = 3;
  const int n_tile_pad = (a.n_tile + 1) / 2 * 2;
  const int work_per_block = n_tile_pad / 2;          // a work item = 256 rows (one 128-row tile per CTA of the pair)
  const int n_workers = (int)gridDim.x / 2;
  const int worker = (int)blockIdx.x / 2;
  const int total_work = work_per_block * a.n_blocks;
  const int n_iter = (total_work + n_workers - 1) / n_workers;

  if (tid == 0) {
    for (int s = 0; s < NWS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); mbar_init(&w_peer[s], 1); }
    for (int s = 0; s < NAS; ++s) { mbar_init(&a_full[s], 2 * R::PROD_W); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * R::EPI_W); }
    for (int s = 0; s < S::NPS; ++s) { mbar_init(&pq_full[s], 1); mbar_init(&pq_empty[s], R::PROD_W); }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc2(tmem_slot, 512);
  if (MODE == MODE_GCL || MODE == MODE_EQUIV) {
    float* vec = reinterpret_cast<float*>(smem + S::OFF_VEC);
    float* wrd = reinterpret_cast<float*>(smem + S::OFF_WRD);
    for (int c = tid; c < H; c += NTHREADS) { vec[c] = a.b2[c]; vec[H + c] = a.w_out[c]; wrd[c] = a.w_rd[c]; wrd[H + c] = a.w_rd[H + c]; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  // everything above touched only weights and on-chip state; results of the previous kernel are read from here on
  pdl_launch_dependents();
  pdl_wait_prior();

  auto tile_of = [&](int iter, int& tile, int& nb, int& row0, int& nrows) {
    const int work = iter * n_workers + worker;
    if (MODE == MODE_GCL || MODE == MODE_EQUIV) {      // edge launches have one column block: no division
      nb = 0;
      tile = 2 * work + (int)crank;
    } else {
      nb = work / work_per_block;
      tile = 2 * (work % work_per_block) + (int)crank;
    }
    if (nb >= a.n_blocks || tile >= a.n_tile) {
      nb = nb >= a.n_blocks ? a.n_blocks - 1 : nb;
      row0 = 0; nrows = 0;
      return;
    }
    row0 = a.tile_row ? a.tile_row[tile] : tile * TM;
    nrows = (a.tile_row ? a.tile_row[tile + 1] : min(a.n_rows, row0 + TM)) - row0;
  };

  // staging decision of a tile (edge modes): all its receivers / senders inside the two TMA boxes
  auto meta_of = [&](int tile, int nrows) -> int4 {
    if ((MODE == MODE_GCL || MODE == MODE_EQUIV) && a.tile_meta != nullptr && nrows > 0) return __ldg(a.tile_meta + tile);
    return make_int4(0, 0, 0, 0);
  };

  if (warp == WARP_LOAD) {
    // =========================== W-stage loader (TMA engine) ==============================================
    if ((MODE == MODE_GCL || MODE == MODE_EQUIV) && lane == 2 && a.tile_meta != nullptr) {
      // staged projections: per k-slab one tensor copy of the tile's receiver rows (P columns) and one of its sender
      // rows (Q columns); rows past the end of the tensor are zero-filled by the TMA unit
      uint32_t pit = 0;
      for (int iter = 0; iter < n_iter; ++iter) {
        int tile, nb, row0, nrows;
        tile_of(iter, tile, nb, row0, nrows);
        const int4 m = meta_of(tile, nrows);
        if (!m.z) continue;
        for (int s = 0; s < H / BK; ++s, ++pit) {
          const int st = pit % NPS;
          mbar_wait(&pq_empty[st], ((pit / NPS) & 1) ^ 1);
          mbar_arrive_expect_tx(&pq_full[st], S::PQ_STAGE);
          uint8_t* dst = smem + S::OFF_PQ + st * S::PQ_STAGE;
          tma_load_2d(dst, &tm_p, s * BK, m.x, &pq_full[st]);
          tma_load_2d(dst + STAGE_PBOX * 256u, &tm_q, H + s * BK, m.y, &pq_full[st]);
        }
      }
    } else if (lane == 0) {
      uint32_t wit = 0;
      for (int iter = 0; iter < n_iter; ++iter) {
        int tile, nb, row0, nrows;
        tile_of(iter, tile, nb, row0, nrows);
        const uint8_t* src = a.w_pack + PACK_HDR + (size_t)nb * a.n_slabs * 2 * S::W_STAGE;
        for (int s = 0; s < a.n_slabs; ++s, ++wit) {
          const int st = wit % NWS;
          mbar_wait(&w_empty[st], ((wit / NWS) & 1) ^ 1);
          mbar_arrive_expect_tx(&w_full[st], S::W_STAGE);
          bulk_g2s(smem + S::OFF_W + st * S::W_STAGE, src + (size_t)(2 * s + crank) * S::W_STAGE, S::W_STAGE, &w_full[st]);
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
  } else if (warp == WARP_MMA) {
    // =========================== MMA issuer (leader CTA) ==================================================
    if (crank == 0) {
      const uint32_t idesc = make_idesc_f16_m256(H);
      uint32_t wit = 0, ait = 0;
      TC_PROF(long long t_acc = 0; long long t_a = 0; long long t_w = 0; const long long t_begin = clock64();)
      for (int iter = 0; iter < n_iter; ++iter) {
        const int region = iter & 1;
        TC_PROF(long long t0 = clock64();)
        mbar_wait_cluster(&acc_empty[region], ((iter >> 1) & 1) ^ 1);
        TC_PROF(t_acc += clock64() - t0;)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + region * 256;
        for (int s = 0; s < a.n_slabs; ++s, ++ait, ++wit) {
          const int ast = (ait / SPS) % NAS, sub = ait % SPS, wst = wit % NWS;
          TC_PROF(t0 = clock64();)
          if (sub == 0) mbar_wait_cluster(&a_full[ast], ((ait / SPS) / NAS) & 1);
          TC_PROF(long long t1 = clock64();)
          mbar_wait(&w_full[wst], (wit / NWS) & 1);
          mbar_wait_cluster(&w_peer[wst], (wit / NWS) & 1);
          TC_PROF(t_a += t1 - t0; t_w += clock64() - t1;)
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_hi = smem_u32(smem + S::OFF_A + ast * S::A_STAGE + sub * S::A_SLAB);
            const uint32_t a_lo = a_hi + TM * 128;
            const uint32_t w_hi = smem_u32(smem + S::OFF_W + wst * S::W_STAGE);
            const uint32_t w_lo = w_hi + S::W_IMG;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t da_hi = make_smem_desc_sw128(a_hi + kk * 32), da_lo = make_smem_desc_sw128(a_lo + kk * 32);
              const uint64_t dw_hi = make_smem_desc_sw128(w_hi + kk * 32), dw_lo = make_smem_desc_sw128(w_lo + kk * 32);
              mma_f16_pair(d_tmem, da_lo, dw_hi, idesc, (s | kk) != 0);
              mma_f16_pair(d_tmem, da_hi, dw_lo, idesc, 1);
              mma_f16_pair(d_tmem, da_hi, dw_hi, idesc, 1);
            }
            mma_commit_pair(&w_empty[wst], cmask);
            if (sub == SPS - 1) mma_commit_pair(&a_empty[ast], cmask);
            if (s == a.n_slabs - 1) mma_commit_pair(&acc_full[region], cmask);
          }
          __syncwarp();
        }
      }
      TC_PROF(if (lane == 0 && blockIdx.x == 0) {
        g_tc16_stats[0] += (unsigned long long)(clock64() - t_begin);
        g_tc16_stats[1] += (unsigned long long)t_acc;
        g_tc16_stats[2] += (unsigned long long)t_a;
        g_tc16_stats[3] += (unsigned long long)t_w;
        g_tc16_stats[4] += 1ull;
        g_tc16_stats[5] += (unsigned long long)n_iter;
      })
    }
  } else if (warp >= R::EPI_W) {
    // =========================== A producers ==============================================================
    // thread (rg, chunk): ROWS_PT CONSECUTIVE tile rows rg * ROWS_PT + p and the 16-byte chunk k = 8 chunk .. +7 of every
    // k-slab.  Edge rows are sorted by (molecule, receiver, sender), so consecutive rows share the receiver: its
    // projection P_i stays in registers across the run and is re-read only where the receiver changes.
    const int pt = tid - EPI_T;
    const int chunk = pt & 7;
    const int rg = pt >> 3;
    uint32_t it = 0, pit = 0;
    float a_mul = 1.0f;
    if (MODE == MODE_RAW && a.a_amax_bits) a_mul = ldexpf(1.0f, amax_exponent(__ldg(a.a_amax_bits)));
    TC_PROF(long long tp_wait = 0; long long tp_comp = 0; long long tp_fence = 0; long long tp_meta = 0;)
    for (int iter = 0; iter < n_iter; ++iter) {
      TC_PROF(const long long tm0 = clock64();)
      int tile, nb, row0, nrows;
      tile_of(iter, tile, nb, row0, nrows);
      const float* pP[ROWS_PT];
      const float* pQ[ROWS_PT];
      uint32_t oP[ROWS_PT], oQ[ROWS_PT];   // staged tiles: byte offsets of the row's P / Q projections inside a stage
      float rr[ROWS_PT], dd[ROWS_PT];
      unsigned lmask = 0;                  // row needs a fresh P load (first row of the thread / receiver changed)
      int prev_i = -1;
      const int4 meta = meta_of(tile, nrows);
      const bool staged = meta.z != 0;
      // rows beyond the end of the tile are CLAMPED to its last row (an empty tile: row 0 of the launch): they produce
      // a copy of a real row, which the epilogue ignores (accumulator rows are independent) -> no validity branches,
      // every load unconditional and in bounds
      const int rlast = nrows > 0 ? nrows - 1 : 0;
      const int rbase0 = nrows > 0 ? row0 : 0;
#pragma unroll
      for (int p = 0; p < ROWS_PT; ++p) {
        const int grow = rbase0 + min(rg * ROWS_PT + p, rlast);
        rr[p] = 0.f; dd[p] = 0.f; pQ[p] = nullptr; oP[p] = 0; oQ[p] = 0;
        if (MODE == MODE_DENSE) {
          pP[p] = a.a1 + (size_t)grow * a.k1 + 4 * chunk;
          pQ[p] = a.a2 ? a.a2 + (size_t)grow * a.k2 + 4 * chunk : nullptr;
        } else {
          const int i = (MODE == MODE_RAW && a.edge_i == nullptr) ? grow : a.edge_i[grow];   // RAW without a row table: A row = tile row
          pP[p] = a.pq + (size_t)i * a.pq_ld + 4 * chunk;
          oP[p] = (uint32_t)(i - meta.x) * 256u + 16u * chunk;
          if (i != prev_i) lmask |= 1u << p;
          prev_i = i;
          if (MODE != MODE_RAW) {
            const int j = a.edge_j[grow];
            pQ[p] = a.pq + (size_t)j * a.pq_ld + H + 4 * chunk;
            oQ[p] = (STAGE_PBOX + (uint32_t)(j - meta.y)) * 256u + 16u * chunk;
            if (a.r_edge) {
              rr[p] = __ldg(a.r_edge + grow);
              dd[p] = __ldg(a.d0_edge + grow);
            } else {
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
      if (MODE == MODE_GCL || MODE == MODE_EQUIV) {
        // the per-row tables of this CTA's NEXT tile (streamed once, so never in L1 by themselves): requested now, they
        // arrive while this tile is produced and the dependent index -> P/Q address chain starts from L1 hits
        const int ntile = 2 * ((iter + 1) * n_workers + worker) + (int)crank;
        const int n2tile = ntile + 2 * n_workers;
        if (a.tile_row && n2tile < a.n_tile && pt == 0) prefetch_l1(a.tile_row + n2tile);   // tile_row two tiles ahead
        if (a.tile_meta && ntile < a.n_tile && pt == 1) prefetch_l1(a.tile_meta + ntile);
        if (iter + 1 < n_iter && ntile < a.n_tile && chunk == 0) {
          const int nrow = (a.tile_row ? a.tile_row[ntile] : ntile * TM) + rg * ROWS_PT;
          if (nrow < a.n_rows) {
            prefetch_l1(a.edge_i + nrow);
            prefetch_l1(a.edge_j + nrow);
            if (a.r_edge) { prefetch_l1(a.r_edge + nrow); prefetch_l1(a.d0_edge + nrow); }
            if (MODE == MODE_EQUIV && a.u_edge) prefetch_l1(a.u_edge + nrow);
          }
        }
      }
      TC_PROF(tp_meta += clock64() - tm0;)
      // edge modes: K = H is a compile-time constant -> the slab loop is unrolled and every P/Q/w_rd load gets an
      // immediate offset (no 64-bit address arithmetic per load)
      constexpr int KS_STATIC = H / BK;
      const int n_slabs_p = (MODE == MODE_DENSE) ? a.n_slabs : KS_STATIC;
      auto run_slabs = [&](auto stg_tag) {
      constexpr bool STG = decltype(stg_tag)::value;   // this tile's projection rows are staged in shared memory (TMA)
#pragma unroll(MODE == MODE_DENSE ? 1 : KS_STATIC)
      for (int s = 0; s < n_slabs_p; ++s, ++it) {
        const int st = (it / SPS) % NAS, sub = it % SPS;
        const int k0 = s * BK;
        uint32_t pqb = 0;
        int pst = 0;
        if constexpr (STG) {
          pst = pit % NPS;
          mbar_wait(&pq_full[pst], (pit / NPS) & 1);
          pqb = sbase + S::OFF_PQ + pst * S::PQ_STAGE;
        }
        auto ld_p = [&](int p, int hsel) -> float4 {
          if constexpr (STG) return lds128f(pqb + oP[p] + hsel * 128);
          else return __ldg(reinterpret_cast<const float4*>(pP[p] + k0 + hsel * KH));
        };
        auto ld_q = [&](int p, int hsel) -> float4 {
          if constexpr (STG) return lds128f(pqb + oQ[p] + hsel * 128);
          else return __ldg(reinterpret_cast<const float4*>(pQ[p] + k0 + hsel * KH));
        };
        TC_PROF(const long long tp0 = clock64(); long long tp1 = tp0;)
        bool waited = sub != 0;                      // the stage was acquired with its first slab
        const uint32_t a_hi = sbase + S::OFF_A + st * S::A_STAGE + sub * S::A_SLAB;
        const uint32_t a_lo = a_hi + TM * 128;
        if constexpr (MODE == MODE_GCL || MODE == MODE_EQUIV) {
          // ---- A = SiLU(P_i + Q_j + w_r r_ij + w_d d0_ij), packed fp32 pairs, shared reciprocals ---------------
          f32x2 WR[4], WD[4];
          {
            const uint32_t s_wrd = sbase + S::OFF_WRD + (k0 + 4 * chunk) * 4;
            const float4 r0 = lds128f(s_wrd), r1 = lds128f(s_wrd + KH * 4);
            const float4 d0 = lds128f(s_wrd + H * 4), d1 = lds128f(s_wrd + H * 4 + KH * 4);
            WR[0] = pk2(r0.x, r0.y); WR[1] = pk2(r0.z, r0.w); WR[2] = pk2(r1.x, r1.y); WR[3] = pk2(r1.z, r1.w);
            WD[0] = pk2(d0.x, d0.y); WD[1] = pk2(d0.z, d0.w); WD[2] = pk2(d1.x, d1.y); WD[3] = pk2(d1.z, d1.w);
          }
          float4 Pc0 = make_float4(0.f, 0.f, 0.f, 0.f), Pc1 = Pc0;
#pragma unroll
          for (int ph = 0; ph < ROWS_PT / 2; ++ph) {   // the Q rows of two tile rows in flight at a time
            float4 q[2][2];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int p = 2 * ph + pp;
              q[pp][0] = ld_q(p, 0);
              q[pp][1] = ld_q(p, 1);
            }
            if (ph == 0) {
              Pc0 = ld_p(0, 0);
              Pc1 = ld_p(0, 1);
            }
            if (!waited) { mbar_wait(&a_empty[st], (((it / SPS) / NAS) & 1) ^ 1); waited = true; TC_PROF(tp1 = clock64();) }
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int p = 2 * ph + pp;
              const int r = rg * ROWS_PT + p;
              if (p > 0 && ((lmask >> p) & 1u)) {
                Pc0 = ld_p(p, 0);
                Pc1 = ld_p(p, 1);
              }
              uint4 hi, lo;
              const f32x2 R2 = pk2(rr[p], rr[p]), D2 = pk2(dd[p], dd[p]);
              f32x2 e0 = fma2(WD[0], D2, fma2(WR[0], R2, add2(pk2(Pc0.x, Pc0.y), pk2(q[pp][0].x, q[pp][0].y))));
              f32x2 e1 = fma2(WD[1], D2, fma2(WR[1], R2, add2(pk2(Pc0.z, Pc0.w), pk2(q[pp][0].z, q[pp][0].w))));
              f32x2 e2 = fma2(WD[2], D2, fma2(WR[2], R2, add2(pk2(Pc1.x, Pc1.y), pk2(q[pp][1].x, q[pp][1].y))));
              f32x2 e3 = fma2(WD[3], D2, fma2(WR[3], R2, add2(pk2(Pc1.z, Pc1.w), pk2(q[pp][1].z, q[pp][1].w))));
              silu_x4(e0, e1);
              silu_x4(e2, e3);
              split_f16x2(e0, hi.x, lo.x); split_f16x2(e1, hi.y, lo.y);
              split_f16x2(e2, hi.z, lo.z); split_f16x2(e3, hi.w, lo.w);
              const uint32_t off = sw128_off(r, chunk);
              sts128(a_hi + off, hi);
              sts128(a_lo + off, lo);
            }
          }
        } else {
          // ---- DENSE / RAW: plain fp32 rows -> fp16 hi | lo ------------------------------------------------------
#pragma unroll
          for (int ph = 0; ph < ROWS_PT / 2; ++ph) {   // two rows at a time: bounds the registers held by loads in flight
            float4 v[2][2], q[2][2];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int p = 2 * ph + pp;
              v[pp][0] = v[pp][1] = q[pp][0] = q[pp][1] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MODE == MODE_DENSE) {
                if (k0 < a.k1) {
                  v[pp][0] = __ldg(reinterpret_cast<const float4*>(pP[p] + k0));
                  v[pp][1] = __ldg(reinterpret_cast<const float4*>(pP[p] + k0 + KH));
                } else {
                  q[pp][0] = __ldg(reinterpret_cast<const float4*>(pQ[p] + (k0 - a.k1)));
                  q[pp][1] = __ldg(reinterpret_cast<const float4*>(pQ[p] + (k0 - a.k1) + KH));
                }
              } else {
                v[pp][0] = __ldg(reinterpret_cast<const float4*>(pP[p] + k0));
                v[pp][1] = __ldg(reinterpret_cast<const float4*>(pP[p] + k0 + KH));
              }
            }
            if (!waited) { mbar_wait(&a_empty[st], (((it / SPS) / NAS) & 1) ^ 1); waited = true; TC_PROF(tp1 = clock64();) }
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int p = 2 * ph + pp;
              const int r = rg * ROWS_PT + p;
              const float vv[8] = {v[pp][0].x, v[pp][0].y, v[pp][0].z, v[pp][0].w, v[pp][1].x, v[pp][1].y, v[pp][1].z, v[pp][1].w};
              const float qq[8] = {q[pp][0].x, q[pp][0].y, q[pp][0].z, q[pp][0].w, q[pp][1].x, q[pp][1].y, q[pp][1].z, q[pp][1].w};
              float e[8];
              if (MODE == MODE_DENSE) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  if (k0 < a.k1) e[c] = vv[c];
                  else e[c] = (a.a2_div != 1.0f) ? __fdiv_rn(qq[c], a.a2_div) : qq[c];
                }
              } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) e[c] = vv[c] * a_mul;
              }
              uint4 hi, lo;
              split_f16x8(e, hi, lo);
              const uint32_t off = sw128_off(r, chunk);
              sts128(a_hi + off, hi);
              sts128(a_lo + off, lo);
            }
          }
        }
        TC_PROF(const long long tp2 = clock64();)
        if constexpr (STG) {                         // every lane's reads of the stage have returned (their values were used)
          __syncwarp();
          if (lane == 0) mbar_arrive(&pq_empty[pst]);
          ++pit;
        }
        if (sub == SPS - 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { if (crank != 0) mbar_arrive_remote(&a_full[st], 0); else mbar_arrive(&a_full[st]); }
        }
        TC_PROF(tp_wait += tp1 - tp0; tp_comp += tp2 - tp1; tp_fence += clock64() - tp2;)
      }
      };
      if constexpr (MODE == MODE_GCL || MODE == MODE_EQUIV) {
        if (staged) run_slabs(std::true_type{}); else run_slabs(std::false_type{});
      } else {
        run_slabs(std::false_type{});
      }
    }
    TC_PROF(if (tid == EPI_T && blockIdx.x == 0) {
      g_tc16_stats[6] += (unsigned long long)tp_wait; g_tc16_stats[7] += (unsigned long long)tp_comp;
      g_tc16_stats[8] += (unsigned long long)tp_fence; g_tc16_stats[9] += (unsigned long long)tp_meta;
    })
  } else {
    // ============= epilogue (warps 0-7: thread = (TMEM lane = row, column half)) ==========================
    const int r = (warp & 3) * 32 + lane;
    const int hf = warp >> 2;               // column half owned by this warp (always 0 when NHALF == 1)
    constexpr int HC = H / NHALF;           // columns per thread
    constexpr int NCH = HC / 32;
    const uint32_t s_i = sbase + S::OFF_SI;          // int   [128] receiver per row
    const uint32_t s_dx = sbase + S::OFF_DX;         // float [128][4]
    const uint32_t s_dot = sbase + S::OFF_DOT;       // float [2][128]
    const uint32_t s_b2 = sbase + S::OFF_VEC + hf * HC * 4;
    const uint32_t s_wo = s_b2 + H * 4;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + hf * HC;
    // 2^-e of the weight images times the round-toward-zero compensation
    float scale = __ldg(reinterpret_cast<const float*>(a.w_pack)) * a.rz_scale;
    if (MODE == MODE_RAW && a.a_amax_bits) scale *= ldexpf(1.0f, -amax_exponent(__ldg(a.a_amax_bits)));
    auto release_acc = [&](int region) {
      __syncwarp();
      if (lane == 0) { if (crank != 0) mbar_arrive_remote(&acc_empty[region], 0); else mbar_arrive(&acc_empty[region]); }
    };
    TC_PROF(long long te_wait = 0; long long te_p1 = 0; long long te_bar = 0; long long te_p2 = 0; long long te_meta = 0;)
    for (int iter = 0; iter < n_iter; ++iter) {
      TC_PROF(const long long te0 = clock64();)
      int tile, nb, row0, nrows;
      tile_of(iter, tile, nb, row0, nrows);
      const bool valid = r < nrows;
      const int region = iter & 1;
      int my_i = -1;
      float ux = 0.f, uy = 0.f, uz = 0.f;
      if (MODE == MODE_GCL || MODE == MODE_EQUIV) {
        if (valid) {
          my_i = a.edge_i[row0 + r];
          if (MODE == MODE_EQUIV && hf == 0) {
            if (a.u_edge) {
              const float4 u4 = __ldg(a.u_edge + row0 + r);
              ux = u4.x; uy = u4.y; uz = u4.z;
            } else {
              EdgeGeom g = edge_geom(a.x, a.x0, my_i, a.edge_j[row0 + r], a.norm_constant);
              ux = g.ux; uy = g.uy; uz = g.uz;
            }
          }
        }
        if (hf == 0) sts32i(s_i + 4 * r, my_i);
      }
      TC_PROF(const long long te1 = clock64();)
      mbar_wait(&acc_full[region], (iter >> 1) & 1);
      tc_fence_after();
      TC_PROF(const long long te2 = clock64(); te_meta += te1 - te0; te_wait += te2 - te1;)
      const uint32_t taddr = tlane + region * 256;

      if (MODE == MODE_DENSE || MODE == MODE_RAW) {
        {
          // lane = accumulator row after tcgen05.ld; a row-per-lane global access would touch 32 lines per instruction, so
          // every 32 x 32 chunk goes through this warp's transposition tile and leaves as 4 rows x 128 contiguous bytes per
          // instruction (lane = (row % 4, 16-byte column group)), with bias / SiLU / residual applied on the way out.
          // The residual rows are requested as one batch of eight coalesced loads per chunk (one exposed L2 latency per
          // chunk, not one per row group), right after the accumulator registers have been handed to shared memory.
          const uint32_t Tw = sbase + S::OFF_T + warp * (32 * 36 * 4);
          const int wrow0 = row0 + (warp & 3) * 32;                 // first global row of this warp's 32 rows
          const int nval = min(32, nrows - (warp & 3) * 32);        // rows of this warp that exist (may be <= 0)
          const int orow_l = lane >> 3, ocol = (lane & 7) * 4;
          const size_t col0 = (size_t)nb * H + hf * HC;
          const bool has_res = (MODE == MODE_DENSE && a.epi == 2);
          const float* bias = (MODE == MODE_DENSE && a.b2) ? a.b2 + col0 : nullptr;
#pragma unroll 1
          for (int cc = 0; cc < NCH; ++cc) {
            uint32_t v[32];
            tmem_ld32(taddr + cc * 32, v);
            tmem_ld_wait();
            if (cc == NCH - 1) { tc_fence_before(); release_acc(region); }     // registers hold the last chunk
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
              sts128f(Tw + (lane * 36 + c4 * 4) * 4, make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]),
                                                                 __uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3])));
            if (has_res) {
              __syncwarp();
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + ocol));
#ifndef GEOLDM_RES_BATCH
#define GEOLDM_RES_BATCH 8
#endif
#pragma unroll
              for (int hb = 0; hb < 8 / GEOLDM_RES_BATCH; ++hb) {          // residual rows in batches of coalesced loads
                float4 rs[GEOLDM_RES_BATCH];
#pragma unroll
                for (int k = 0; k < GEOLDM_RES_BATCH; ++k) {
                  const int rl = max(0, min((hb * GEOLDM_RES_BATCH + k) * 4 + orow_l, nval - 1));   // clamped: in bounds
                  rs[k] = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)(wrow0 + rl) * a.ldo + col0 + cc * 32 + ocol));
                }
#pragma unroll
                for (int k = 0; k < GEOLDM_RES_BATCH; ++k) {
                  const int rl = (hb * GEOLDM_RES_BATCH + k) * 4 + orow_l;
                  const float4 t = lds128f(Tw + (rl * 36 + ocol) * 4);
                  if (rl < nval) {
                    const size_t goff = (size_t)(wrow0 + rl) * a.ldo + col0 + cc * 32 + ocol;
                    *reinterpret_cast<float4*>(a.out + goff) =
                        make_float4(fmaf(t.x, scale, b4.x) + rs[k].x, fmaf(t.y, scale, b4.y) + rs[k].y,
                                    fmaf(t.z, scale, b4.z) + rs[k].z, fmaf(t.w, scale, b4.w) + rs[k].w);
                    if (a.zero_buf) *reinterpret_cast<float4*>(a.zero_buf + goff) = make_float4(0.f, 0.f, 0.f, 0.f);
                  }
                }
              }
            } else {
              __syncwarp();
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + ocol));
#pragma unroll
              for (int it8 = 0; it8 < 8; ++it8) {
                const int rl = it8 * 4 + orow_l;
                if (rl < nval) {
                  const float4 t = lds128f(Tw + (rl * 36 + ocol) * 4);
                  float o[4] = {fmaf(t.x, scale, b4.x), fmaf(t.y, scale, b4.y), fmaf(t.z, scale, b4.z), fmaf(t.w, scale, b4.w)};
                  if (MODE == MODE_DENSE && a.epi == 1) { o[0] = silu(o[0]); o[1] = silu(o[1]); o[2] = silu(o[2]); o[3] = silu(o[3]); }
                  *reinterpret_cast<float4*>(a.out + (size_t)(wrow0 + rl) * a.ldo + col0 + cc * 32 + ocol) =
                      make_float4(o[0], o[1], o[2], o[3]);
                }
              }
            }
            __syncwarp();
          }
        }
      } else {
        // ---- pass 1: m = SiLU(scale * D + b2), partial row dot with w_att / w6 over this thread's column half -----
        // (packed fp32 pairs, one reciprocal shared by two values, two partial dots)
        f32x2 dot2 = pk2(0.f, 0.f);
        const f32x2 scale2 = pk2(scale, scale);
        // one 32-column chunk: bias, SiLU, partial dot; the SiLU values go back into v (GCL writes them to TMEM for pass 2)
        auto p1_chunk = [&](uint32_t (&v)[32], int cc) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 b4 = lds128f(s_b2 + (cc * 32 + c4 * 4) * 4);
            const float4 w4 = lds128f(s_wo + (cc * 32 + c4 * 4) * 4);
            f32x2 m0 = fma2(pk2(__uint_as_float(v[c4 * 4 + 0]), __uint_as_float(v[c4 * 4 + 1])), scale2, pk2(b4.x, b4.y));
            f32x2 m1 = fma2(pk2(__uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3])), scale2, pk2(b4.z, b4.w));
            silu_x4(m0, m1);
            dot2 = fma2(pk2(w4.x, w4.y), m0, dot2);
            dot2 = fma2(pk2(w4.z, w4.w), m1, dot2);
            float f0, f1, f2, f3;
            upk2(m0, f0, f1);
            upk2(m1, f2, f3);
            v[c4 * 4 + 0] = __float_as_uint(f0); v[c4 * 4 + 1] = __float_as_uint(f1);
            v[c4 * 4 + 2] = __float_as_uint(f2); v[c4 * 4 + 3] = __float_as_uint(f3);
          }
          if (MODE == MODE_GCL) tmem_st32(taddr + cc * 32, v);
        };
#ifndef GEOLDM_P1_PIPE
#define GEOLDM_P1_PIPE 1
#endif
        if constexpr (GEOLDM_P1_PIPE && NCH % 2 == 0 && MODE == MODE_GCL) {
          // two register buffers: the TMEM load of chunk cc + 1 is in flight while chunk cc is processed
          uint32_t va[32], vb[32];
          tmem_ld32(taddr, va);
#pragma unroll
          for (int cc = 0; cc < NCH; cc += 2) {
            tmem_ld_wait();
            tmem_ld32(taddr + (cc + 1) * 32, vb);
            p1_chunk(va, cc);
            tmem_ld_wait();
            if (cc + 2 < NCH) tmem_ld32(taddr + (cc + 2) * 32, va);
            p1_chunk(vb, cc + 1);
          }
        } else {
#pragma unroll 1
          for (int cc = 0; cc < NCH; ++cc) {
            uint32_t v[32];
            tmem_ld32(taddr + cc * 32, v);
            tmem_ld_wait();
            p1_chunk(v, cc);
          }
        }
        float dot;
        {
          float d0, d1;
          upk2(dot2, d0, d1);
          dot = d0 + d1;
        }
        TC_PROF(const long long te3 = clock64(); te_p1 += te3 - te2;)
        float full_dot = dot;
        if (NHALF == 2) {                       // the two warps sharing this lane quarter exchange their partial dots
          sts32f(s_dot + (hf * TM + r) * 4, dot);
          if (MODE == MODE_GCL) tmem_st_wait();
          named_bar_sync(2 + (warp & 3), 64);
          full_dot = lds32f(s_dot + r * 4) + lds32f(s_dot + (TM + r) * 4);
        } else {
          __syncwarp();                         // s_i of this warp's rows is visible to its lanes
        }
        const int prev_i = __shfl_up_sync(0xffffffffu, my_i, 1);
        const bool head = valid && (lane == 0 || prev_i != my_i);
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        const int nval = __popc(__ballot_sync(0xffffffffu, valid));
        if (MODE == MODE_EQUIV) {
          tc_fence_before();
          release_acc(region);
          if (hf == 0) {
            float phi = a.use_tanh ? tanhf(full_dot) : full_dot;
            float dx = __fmul_rn(ux, phi), dy = __fmul_rn(uy, phi), dz = __fmul_rn(uz, phi);
            if (a.use_tanh) { dx = __fmul_rn(dx, a.coords_range); dy = __fmul_rn(dy, a.coords_range); dz = __fmul_rn(dz, a.coords_range); }
            sts128f(s_dx + 16 * r, make_float4(valid ? dx : 0.f, valid ? dy : 0.f, valid ? dz : 0.f, 0.f));
            __syncwarp();
            if (head) {
              const unsigned after = hm & ~((2u << lane) - 1u);
              const int q1 = after ? (__ffs(after) - 1) : nval;
              float sx = 0.f, sy = 0.f, sz = 0.f;
              for (int q = lane; q < q1; ++q) {
                const float4 d = lds128f(s_dx + 16 * ((warp & 3) * 32 + q));
                sx += d.x; sy += d.y; sz += d.z;
              }
              atomicAdd(a.out + (size_t)my_i * 3, sx);
              atomicAdd(a.out + (size_t)my_i * 3 + 1, sy);
              atomicAdd(a.out + (size_t)my_i * 3 + 2, sz);
            }
            __syncwarp();
          }
          if (NHALF == 2) named_bar_sync(2 + (warp & 3), 64);
        } else {
          float g = a.attention ? sigmoidf_(full_dot + __ldg(a.b_out)) : 1.0f;
          if (!valid) g = 0.f;
          const f32x2 g2 = pk2(g, g);
          TC_PROF(const long long te4 = clock64(); te_bar += te4 - te3;)
          // ---- pass 2: e = m * gate, summed over each receiver run of this warp's 32 rows -------------------------
          // 32 x 32 chunks go through this warp's [32][36] fp32 transposition tile.  Static path: lane (grp, c4) owns the
          // 8-row group grp and 4 columns, multiplies by the row gates inside an unrolled 8-step sum, exchanges the partial
          // sum of a run that continues into the next group through the tile's pad columns and emits one 16-byte vector
          // reduction per (run, 4 columns).  Generic path (runs shorter than 8 rows): lane (half, cp) sums columns 2 cp,
          // 2 cp + 1 over every second row of a run.  The summation order is fixed in both.
          const uint32_t Tw = sbase + S::OFF_T + warp * (32 * 36 * 4);
          // Run structure of this warp's 32 rows, seen by lane (grp, c4): row group [8 grp, 8 grp + 8) x columns 4 c4 .. +3 of
          // a chunk.  The first invalid row closes the last run (its "receiver" is -1: nothing is emitted for it).
          const int grp = lane >> 3, c4 = lane & 7;
          const unsigned hg = ((hm | (nval < 32 ? (1u << nval) : 0u)) >> (8 * grp)) & 0xFFu;
          const unsigned inner = hg & 0xFEu;              // run boundaries strictly inside the group
          // static path: at most one boundary inside every 8-row group (always true when runs are >= 8 rows long, i.e.
          // molecules of >= 9 atoms); shorter runs take the generic loop below
          const bool fast = __all_sync(0xffffffffu, __popc(inner) <= 1);
          uint32_t v[32];
          tmem_ld32(taddr, v);
          if (fast) {
            const int b = inner ? (__ffs(inner) - 1) : 8;                  // rows q < b: "lo" segment, q >= b: "hi" segment
            const bool has_inner = inner != 0u;
            const bool first_head = (hg & 1u) != 0u || grp == 0;           // the group's first row opens a run (of this warp)
            const unsigned fhm = __ballot_sync(0xffffffffu, first_head);
            const unsigned psm = __ballot_sync(0xffffffffu, !has_inner && !first_head);   // group lies inside one older run
            const bool end_emit = grp == 3 || ((fhm >> (8 * (grp + 1))) & 1u) != 0u;     // the run open at the group's end closes
            // carry-in = rows of the run containing row 8 grp that sit in earlier groups of this warp
            const bool a1 = !first_head;
            const bool a2 = a1 && grp >= 2 && ((psm >> (8 * (grp - 1))) & 1u) != 0u;
            const bool a3 = a2 && grp >= 3 && ((psm >> (8 * (grp - 2))) & 1u) != 0u;
            const int recv_lo = __shfl_sync(0xffffffffu, my_i, 8 * grp);
            const int recv_hi = __shfl_sync(0xffffffffu, my_i, 8 * grp + 7);
            const bool emit_lo = (has_inner || end_emit) && recv_lo >= 0;
            const bool emit_hi = has_inner && end_emit && recv_hi >= 0;
            float* const out_lo = a.out + (size_t)(recv_lo < 0 ? 0 : recv_lo) * H + hf * HC + 4 * c4;
            float* const out_hi = a.out + (size_t)(recv_hi < 0 ? 0 : recv_hi) * H + hf * HC + 4 * c4;
            const uint32_t tp = Tw + ((8 * grp) * 36 + 4 * c4) * 4;
            const uint32_t tv = Tw + (lane * 36 + 32) * 4;                 // pad columns of row `lane`: carry exchange
            // the gate is applied inside the sums (fma with the row's gate) instead of on every element before the
            // transposition: the 8 gates of this lane's row group, through a warp-private shared-memory array
            const uint32_t s_g = sbase + S::OFF_PS + warp * (32 * 4);
            sts32f(s_g + 4 * lane, g);
            __syncwarp();
            const float4 ga = lds128f(s_g + 32 * grp), gb = lds128f(s_g + 32 * grp + 16);
            const f32x2 gq[8] = {pk2(ga.x, ga.x), pk2(ga.y, ga.y), pk2(ga.z, ga.z), pk2(ga.w, ga.w),
                                 pk2(gb.x, gb.x), pk2(gb.y, gb.y), pk2(gb.z, gb.z), pk2(gb.w, gb.w)};
#pragma unroll 1
            for (int cc = 0; cc < NCH; ++cc) {
              tmem_ld_wait();
              if (cc == NCH - 1) { tc_fence_before(); release_acc(region); }
#pragma unroll
              for (int c4i = 0; c4i < 8; ++c4i)
                sts128(Tw + (lane * 36 + c4i * 4) * 4, make_uint4(v[c4i * 4], v[c4i * 4 + 1], v[c4i * 4 + 2], v[c4i * 4 + 3]));
              if (cc + 1 < NCH) tmem_ld32(taddr + (cc + 1) * 32, v);   // in flight while the rows of this chunk are summed
              __syncwarp();
              f32x2 lo0 = pk2(0.f, 0.f), lo1 = lo0, hi0 = lo0, hi1 = lo0;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                f32x2 t0, t1;
                lds128p(tp + q * (36 * 4), t0, t1);
                if (q < b) { lo0 = fma2(gq[q], t0, lo0); lo1 = fma2(gq[q], t1, lo1); }
                else { hi0 = fma2(gq[q], t0, hi0); hi1 = fma2(gq[q], t1, hi1); }
              }
              // the segment still open at the end of the group, offered to the following groups
              sts128p(tv, has_inner ? hi0 : lo0, has_inner ? hi1 : lo1);
              __syncwarp();
              f32x2 c0 = pk2(0.f, 0.f), c1 = c0;
              if (a3) { f32x2 t0, t1; lds128p(tv - 24 * (36 * 4), t0, t1); c0 = t0; c1 = t1; }
              if (a2) { f32x2 t0, t1; lds128p(tv - 16 * (36 * 4), t0, t1); c0 = add2(c0, t0); c1 = add2(c1, t1); }
              if (a1) { f32x2 t0, t1; lds128p(tv - 8 * (36 * 4), t0, t1); c0 = add2(c0, t0); c1 = add2(c1, t1); }
              if (emit_lo) red_add_v4p(out_lo + cc * 32, add2(c0, lo0), add2(c1, lo1));
              if (emit_hi) red_add_v4p(out_hi + cc * 32, hi0, hi1);
              __syncwarp();
            }
          } else {
          const int half = lane >> 4, cp = lane & 15;
#pragma unroll 1
          for (int cc = 0; cc < NCH; ++cc) {
            tmem_ld_wait();
            if (cc == NCH - 1) { tc_fence_before(); release_acc(region); }
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              float e0, e1, e2, e3;
              upk2(mul2(pk2(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1])), g2), e0, e1);
              upk2(mul2(pk2(__uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3])), g2), e2, e3);
              sts128f(Tw + (lane * 36 + c4 * 4) * 4, make_float4(e0, e1, e2, e3));
            }
            if (cc + 1 < NCH) tmem_ld32(taddr + (cc + 1) * 32, v);   // in flight while the rows of this chunk are summed
            __syncwarp();
            unsigned rest = hm;
            while (rest) {
              const int q0 = __ffs(rest) - 1;
              rest &= rest - 1;
              const int q1 = rest ? (__ffs(rest) - 1) : nval;
              const int pi = __shfl_sync(0xffffffffu, my_i, q0);
              f32x2 sA = pk2(0.f, 0.f), sB = sA;
              // rows q0 + half, + 2, ... of the run, four 64-bit loads in flight (rows past the run read as zero)
              for (int q = q0 + half; q < q1; q += 8) {
                const uint32_t tp = Tw + (q * 36 + 2 * cp) * 4;
                const f32x2 t0 = lds64(tp);
                const f32x2 t1 = (q + 2 < q1) ? lds64(tp + 2 * 36 * 4) : pk2(0.f, 0.f);
                const f32x2 t2 = (q + 4 < q1) ? lds64(tp + 4 * 36 * 4) : pk2(0.f, 0.f);
                const f32x2 t3 = (q + 6 < q1) ? lds64(tp + 6 * 36 * 4) : pk2(0.f, 0.f);
                sA = add2(sA, add2(t0, t2));
                sB = add2(sB, add2(t1, t3));
              }
              sA = add2(sA, sB);
              float s0, s1;
              upk2(sA, s0, s1);
              s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
              s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
              if (half == 0) red_add_v2(a.out + (size_t)pi * H + hf * HC + cc * 32 + 2 * cp, s0, s1);
            }
            __syncwarp();
          }
          }
          TC_PROF(te_p2 += clock64() - te4;)
          named_bar_sync(2 + (warp & 3), 64);
        }
      }
    }
    TC_PROF(if (tid == 0 && blockIdx.x == 0) {
      g_tc16_stats[10] += (unsigned long long)te_wait; g_tc16_stats[11] += (unsigned long long)te_p1;
      g_tc16_stats[12] += (unsigned long long)te_bar; g_tc16_stats[13] += (unsigned long long)te_p2;
      g_tc16_stats[14] += (unsigned long long)te_meta;
    })
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
  TC_PROF(if (tid == 0 && blockIdx.x == 0) g_tc16_stats[15] += (unsigned long long)(clock64() - t_entry);)
}

// ---- TMA tensor maps over the [n_node][2H] projection matrix (row pitch pq_ld floats): boxes of 64 columns x box_rows rows.
// cuTensorMapEncodeTiled is a host-only driver call (no device work); the maps of the last few (pointer, shape) combinations
// are kept, a CUDA graph bakes them into its kernel nodes.
struct StageMaps { CUtensorMap p, q; };
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_stage_maps(const float* pq, int pq_ld, int n_node, int H, StageMaps& out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("tc16: cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
      return -2;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  struct Key { const float* pq; int ld, n, H; };
  static thread_local Key keys[8];
  static thread_local StageMaps maps[8];
  static thread_local int next = 0;
  for (int k = 0; k < 8; ++k)
    if (keys[k].pq == pq && keys[k].ld == pq_ld && keys[k].n == n_node && keys[k].H == H) { out = maps[k]; return 0; }
  const cuuint64_t dims[2] = {(cuuint64_t)(2 * H), (cuuint64_t)n_node};
  const cuuint64_t strides[1] = {(cuuint64_t)pq_ld * sizeof(float)};
  const cuuint32_t estr[2] = {1, 1};
  const cuuint32_t box_p[2] = {(cuuint32_t)BK, STAGE_PBOX}, box_q[2] = {(cuuint32_t)BK, STAGE_QBOX};
  StageMaps m;
  CUresult r1 = encode(&m.p, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pq), dims, strides, box_p, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = encode(&m.q, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pq), dims, strides, box_q, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    set_error("tc16: cuTensorMapEncodeTiled failed (%d, %d) for pq=%p ld=%d n=%d", (int)r1, (int)r2, (const void*)pq, pq_ld, n_node);
    return -2;
  }
  keys[next] = Key{pq, pq_ld, n_node, H};
  maps[next] = m;
  next = (next + 1) % 8;
  out = m;
  return 0;
}

template <int H, int MODE>
int launch_mode(const Args& a, cudaStream_t st, const StageMaps* sm = nullptr) {
  using S = Smem<H, MODE>;
  static DeviceOnce once;
  bool fresh;
  const int slot = device_slot(once, fresh);
  if (fresh) {
    cudaError_t e = cudaFuncSetAttribute(tc16_kernel<H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::ALLOC);
    if (e != cudaSuccess) {
      set_error("tc16_kernel: cannot reserve %u bytes of shared memory: %s", S::ALLOC, cudaGetErrorString(e));
      return -2;
    }
    once.done[slot] = true;
  }
  const int sm_count = once.sm_count[slot];
  if (a.n_tile == 0) return 0;
  const int work = (a.n_tile + 1) / 2 * 2 * a.n_blocks;
  int grid = work < sm_count ? work : sm_count;
  grid = grid / 2 * 2;
  Args args = a;
  args.rz_scale = 1.0f + RZ_BIAS_PER_MMA * (float)(a.n_slabs * (BK / 16) * 3);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Roles<MODE>::NTHREADS);
  cfg.dynamicSmemBytes = S::ALLOC;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("GEOLDM_TC_PDL"); pdl = e ? atoi(e) : 1; }
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  static const StageMaps no_maps{};
  if (!sm) { sm = &no_maps; args.tile_meta = nullptr; }
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc16_kernel<H, MODE>, args, sm->p, sm->q);
  if (e != cudaSuccess) {
    set_error("tc16_kernel launch: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

template <int MODE>
int launch_h(int H, const Args& a, cudaStream_t st, const StageMaps* sm = nullptr) {
  switch (H) {
    case 64: return launch_mode<64, MODE>(a, st, sm);
    case 128: return launch_mode<128, MODE>(a, st, sm);
    case 192: return launch_mode<192, MODE>(a, st, sm);
    case 256: return launch_mode<256, MODE>(a, st, sm);
    default: set_error("tcgen05 kernels support hidden_nf 64/128/192/256, got %d", H); return -1;
  }
}

// fp16 split of the scaled weights, one thread per element: W[n][k] -> block n/H, slab k/64, N-half, hi|lo images
__global__ void pack16_amax_kernel(const float* __restrict__ w, size_t n, unsigned* __restrict__ amax_bits) {
  float m = 0.f;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(w) & 15u) == 0) {             // 16-byte loads over the aligned bulk, scalars for the tail
    const size_t n4 = n / 4;
    for (size_t i = tid; i < n4; i += nthr) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (size_t i = 4 * n4 + tid; i < n; i += nthr) m = fmaxf(m, fabsf(w[i]));
  } else {
    for (size_t i = tid; i < n; i += nthr) m = fmaxf(m, fabsf(w[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // one atomic per block (every warp on the same word serialises in L2: ~10 k atomics for an [E, H] operand before)
  __shared__ float wmax[32];
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned k = 1; k < (blockDim.x + 31) / 32; ++k) m = fmaxf(m, wmax[k]);
    atomicMax(amax_bits, __float_as_uint(m));                    // non-negative floats order like uints
  }
}

// one element (output row n, input column kk) of W, already scaled, into the hi | lo images of a pack
__device__ __forceinline__ void pack16_store(int H, int k, int n, int kk, float v, uint8_t* __restrict__ pack) {
  const int nb = n / H, nl = n % H, slab = kk / 64, ko = kk % 64;
  const int kl = 8 * ((ko & 31) >> 2) + (ko & 3) + 4 * (ko >> 5);   // position inside the slab after the K permutation
  const int n_slabs = k / 64;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int NH = H / 2, half = nl / NH, rl = nl % NH;
  // (row rl, 16-byte chunk kl/8) of a SWIZZLE_128B K-major image with 128-byte rows, element index inside the chunk kl%8
  const size_t off = (size_t)(rl >> 3) * 1024 + (rl & 7) * 128 + ((((kl >> 3) ^ (rl & 7)) << 4) | ((kl & 7) << 1));
  uint8_t* img = pack + PACK_HDR + (size_t)((nb * n_slabs + slab) * 2 + half) * 2 * (size_t)NH * 128;
  *reinterpret_cast<__half*>(img + off) = hi;
  *reinterpret_cast<__half*>(img + (size_t)NH * 128 + off) = lo;
}
// scale exponent e with max|w| 2^e in [2^13, 2^14)  (zero matrix: e = 0)
__device__ __forceinline__ int pack16_exponent(float amax) {
  int e = 0;
  if (amax > 0.f) { int ex; frexpf(amax, &ex); e = 14 - ex; }       // amax = f * 2^ex, f in [0.5, 1)
  return max(-100, min(100, e));
}
__global__ void pack16_kernel(int H, const float* __restrict__ w, int n_out, int k, uint8_t* __restrict__ pack, int transposed) {
  const int e = pack16_exponent(__uint_as_float(*reinterpret_cast<const unsigned*>(pack + 4)));
  const float scale = ldexpf(1.0f, e);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *reinterpret_cast<float*>(pack) = ldexpf(1.0f, -e);
  if (idx >= (size_t)n_out * k) return;
  const int n = (int)(idx / k), kk = (int)(idx % k);
  // transposed: the source is [k][n_out] row-major (a Linear weight read as the operand of its input-gradient GEMM)
  const float v = (transposed ? w[(size_t)kk * n_out + n] : w[idx]) * scale;   // exact (power of two)
  pack16_store(H, k, n, kk, v, pack);
}
// Small square weights (training: they change every optimiser step and are packed once per step and direction): ONE
// launch finds max|w| (every block on its own: the matrix is 64-256 KB of L2-resident data), writes the headers and its
// share of both images - the forward operand (rows = output features) and the operand of the input-gradient GEMM (rows =
// input features) - instead of memset + amax + pack launches per image.  (The image stores are scattered 2-byte writes: a
// single block needs 100 us for them, 36 blocks 20 us; hence one element per thread, like pack16_kernel.)
constexpr int PAIR_THREADS = 256;
__global__ void __launch_bounds__(PAIR_THREADS) pack16_pair_kernel(int H, const float* __restrict__ w, uint8_t* __restrict__ pack_fwd,
                                                                   uint8_t* __restrict__ pack_t) {
  __shared__ float wmax[32];
  __shared__ float s_amax;
  const int tot = H * H;                          // a multiple of 4096
  float m = 0.f;
  if ((reinterpret_cast<uintptr_t>(w) & 15u) == 0) {
    for (int i = threadIdx.x; i < tot / 4; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
  } else {
    for (int i = threadIdx.x; i < tot; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned q = 1; q < (blockDim.x + 31) / 32; ++q) m = fmaxf(m, wmax[q]);
    s_amax = m;
  }
  __syncthreads();
  const int e = pack16_exponent(s_amax);
  const float scale = ldexpf(1.0f, e);
  if (blockIdx.x == 0 && threadIdx.x < 2) {
    uint8_t* p = threadIdx.x == 0 ? pack_fwd : pack_t;
    if (p) { *reinterpret_cast<float*>(p) = ldexpf(1.0f, -e); *reinterpret_cast<unsigned*>(p + 4) = __float_as_uint(s_amax); }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
    const int n = i / H, kk = i % H;
    const float v = w[i] * scale;                                   // exact (power of two)
    if (pack_fwd) pack16_store(H, H, n, kk, v, pack_fwd);
    if (pack_t) pack16_store(H, H, kk, n, v, pack_t);
  }
}
// =====================================================================================================================
// Fused node chain of one GCL (egnn/egnn_new.py:47-56) plus the first-layer projections that read its output:
//     t1 = SiLU([h | agg / div] W1^T + b1)         (node_mlp.0, K = 2H)
//     h' = h + t1 W2^T + b2                        (node_mlp.2 + residual, K = H)
//     pq = h' Wp^T + bp                            (n_pb column blocks of H: the P|Q projections of the next edge MLPs)
// ONE persistent launch per GCL instead of three dense launches: a CTA pair keeps a 256-row tile on chip through all three
// GEMMs.  t1 and h' never exist as fp32 matrices between the GEMMs: the epilogue of a phase writes its result straight into
// shared memory as the NEXT phase's A operand (fp16 hi | lo images in the canonical SWIZZLE_128B K-major layout, same
// K order inside a slab as the producers use), so phases 2 and 3 need no producer work, no global round trip and no
// extra launch.  Phase 1 streams [h | agg] through the usual producer ring, which aliases the first two slabs of the image
// (the images are written only after the phase-1 MMAs have completed; the producers of the next tile wait for img_free).
// Accumulator regions alternate per GEMM (phase 1 -> 0, phase 2 -> 1, phase 3 blocks 0,1,0,1), so an epilogue overlaps the
// MMAs that follow it.  Roles: 8 epilogue warps, 16 producer warps, loader, MMA issuer (832 threads).
struct ChainArgs {
  int n_tile, n_rows;
  const float* h; const float* agg; float agg_div;
  const uint8_t* pack1; const float* b1;
  const uint8_t* pack2; const float* b2;
  const uint8_t* pack3; const float* b3; int n_pb;
  float* h_out; float* pq_out; float* zero_buf;
  float rz1, rz2;                 // round-toward-zero compensation for K = 2H and K = H
};

template <int H>
struct ChainSmem {
  static constexpr int KS = H / BK;
  static constexpr int NWS = 2, NAS = 2;
  static constexpr int IMG_SLABS = KS > NAS ? KS : NAS;
  static constexpr uint32_t NH = H / 2;
  static constexpr uint32_t W_IMG = NH * 128u;
  static constexpr uint32_t W_STAGE = 2u * W_IMG;
  static constexpr uint32_t A_SLAB = 2u * TM * 128u;
  static constexpr uint32_t OFF_W = 0;
  static constexpr uint32_t OFF_A = OFF_W + NWS * W_STAGE;           // producer ring = image slabs 0, 1
  static constexpr uint32_t OFF_T = OFF_A + IMG_SLABS * A_SLAB;       // 8 warp-private [32][20] fp32 transposition tiles
  static constexpr uint32_t OFF_BAR = OFF_T + 8u * 32 * 20 * 4;
  static constexpr int N_BAR = 3 * NWS + 2 * NAS + 4 + KS + 1;
  static constexpr uint32_t OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr uint32_t BYTES = OFF_TMEM + 16;
  static constexpr uint32_t ALLOC = BYTES + 1024;
};

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

#ifndef GEOLDM_CHAIN_PROD_W
#define GEOLDM_CHAIN_PROD_W 16
#endif
constexpr int CHAIN_THREADS = 32 * (8 + GEOLDM_CHAIN_PROD_W + 2);
template <int H>
__global__ void __launch_bounds__(CHAIN_THREADS, 1) chain16_kernel(const ChainArgs a) {
  using S = ChainSmem<H>;
  constexpr int KS = S::KS, NWS = S::NWS, NAS = S::NAS;
  constexpr int EPI_W = 8, PROD_W = GEOLDM_CHAIN_PROD_W, EPI_T = 32 * EPI_W, WARP_LOAD = EPI_W + PROD_W, WARP_MMA = WARP_LOAD + 1;
  constexpr int ROWS_PT = 128 / (4 * PROD_W);     // 2 (16 producer warps) or 4 (8)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NWS;
  uint64_t* w_peer = bars + 2 * NWS;
  uint64_t* a_full = bars + 3 * NWS;
  uint64_t* a_empty = a_full + NAS;
  uint64_t* acc_full = a_empty + NAS;      // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* img_full = acc_empty + 2;      // [KS] slab s of the operand image is complete in BOTH CTAs
  uint64_t* img_free = img_full + KS;      // [1] every MMA of the tile has completed: ring / image may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  constexpr uint16_t cmask = 3;
  const int n_pairs = (a.n_tile + 1) / 2;
  const int n_workers = (int)gridDim.x / 2;
  const int worker = (int)blockIdx.x / 2;
  const int n_iter = worker < n_pairs ? (n_pairs - worker + n_workers - 1) / n_workers : 0;   // same for both CTAs of a pair

  if (tid == 0) {
    for (int s = 0; s < NWS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); mbar_init(&w_peer[s], 1); }
    for (int s = 0; s < NAS; ++s) { mbar_init(&a_full[s], 2 * PROD_W); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * EPI_W); }
    for (int s = 0; s < KS; ++s) mbar_init(&img_full[s], 2 * EPI_W);       // (4 lane quarters x 2 column pieces) x 2 CTAs
    mbar_init(img_free, 1);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  pdl_launch_dependents();
  pdl_wait_prior();

  auto tile_of = [&](int iter, int& row0, int& nrows) {
    const int tile = 2 * (iter * n_workers + worker) + (int)crank;
    row0 = tile * TM;
    nrows = tile < a.n_tile ? min(TM, a.n_rows - row0) : 0;
    if (nrows <= 0) { row0 = 0; nrows = 0; }
  };
  const int stages_per_iter = (3 + a.n_pb) * KS;

  if (warp == WARP_LOAD) {
    if (lane == 0) {
      uint32_t wit = 0;
      for (int iter = 0; iter < n_iter; ++iter)
        for (int ph = 0; ph < 3; ++ph) {
          const uint8_t* pack = (ph == 0 ? a.pack1 : ph == 1 ? a.pack2 : a.pack3) + PACK_HDR;
          const int nb = ph == 2 ? a.n_pb : 1, ns = ph == 0 ? 2 * KS : KS;
          for (int bs = 0; bs < nb * ns; ++bs, ++wit) {
            const int st = wit % NWS;
            mbar_wait(&w_empty[st], ((wit / NWS) & 1) ^ 1);
            mbar_arrive_expect_tx(&w_full[st], S::W_STAGE);
            bulk_g2s(smem + S::OFF_W + st * S::W_STAGE, pack + (size_t)(2 * bs + crank) * S::W_STAGE, S::W_STAGE, &w_full[st]);
          }
        }
    } else if (lane == 1 && crank == 1) {
      uint32_t wit = 0;
      for (int iter = 0; iter < n_iter; ++iter)
        for (int s = 0; s < stages_per_iter; ++s, ++wit) {
          const int st = wit % NWS;
          mbar_wait(&w_full[st], (wit / NWS) & 1);
          mbar_arrive_remote(&w_peer[st], 0);
        }
    }
  } else if (warp == WARP_MMA) {
    if (crank == 0) {
      const uint32_t idesc = make_idesc_f16_m256(H);
      uint32_t wit = 0, ait = 0, acc_it = 0, img_gen = 0;
      // one k-slab: A images at a_hi (lo image TM * 128 bytes behind), W stage wst
      TC_PROF(long long t_acc = 0; long long t_a = 0; long long t_w = 0; long long t_i2 = 0; long long t_i3 = 0; long long t_p1 = 0;
              long long t_p2 = 0; long long t_p3 = 0; const long long t_begin = clock64();)
      auto slab = [&](uint32_t a_hi, int wst, bool first, uint32_t d_tmem) {
        TC_PROF(const long long tw0 = clock64();)
        mbar_wait(&w_full[wst], (wit / NWS) & 1);
        mbar_wait_cluster(&w_peer[wst], (wit / NWS) & 1);
        TC_PROF(t_w += clock64() - tw0;)
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_lo = a_hi + TM * 128;
          const uint32_t w_hi = smem_u32(smem + S::OFF_W + wst * S::W_STAGE);
          const uint32_t w_lo = w_hi + S::W_IMG;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da_hi = make_smem_desc_sw128(a_hi + kk * 32), da_lo = make_smem_desc_sw128(a_lo + kk * 32);
            const uint64_t dw_hi = make_smem_desc_sw128(w_hi + kk * 32), dw_lo = make_smem_desc_sw128(w_lo + kk * 32);
            mma_f16_pair(d_tmem, da_lo, dw_hi, idesc, !(first && kk == 0));
            mma_f16_pair(d_tmem, da_hi, dw_lo, idesc, 1);
            mma_f16_pair(d_tmem, da_hi, dw_hi, idesc, 1);
          }
          mma_commit_pair(&w_empty[wst], cmask);
        }
      };
      for (int iter = 0; iter < n_iter; ++iter) {
        // ---- phase 1: [h | agg] from the producer ring ----------------------------------------------------------------
        TC_PROF(const long long tp1 = clock64();)
        {
          const int region = acc_it & 1;
          TC_PROF(const long long ta0 = clock64();)
          mbar_wait_cluster(&acc_empty[region], ((acc_it >> 1) & 1) ^ 1);
          TC_PROF(t_acc += clock64() - ta0;)
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + region * 256;
          for (int s = 0; s < 2 * KS; ++s, ++ait, ++wit) {
            const int ast = ait % NAS;
            TC_PROF(const long long tf0 = clock64();)
            mbar_wait_cluster(&a_full[ast], (ait / NAS) & 1);
            TC_PROF(t_a += clock64() - tf0;)
            slab(sbase + S::OFF_A + ast * S::A_SLAB, wit % NWS, s == 0, d_tmem);
            if (lane == 0) {
              mma_commit_pair(&a_empty[ast], cmask);
              if (s == 2 * KS - 1) mma_commit_pair(&acc_full[region], cmask);
            }
            __syncwarp();
          }
          ++acc_it;
        }
        // ---- phase 2 (A = t1 image) and phase 3 (A = h' image, n_pb column blocks) ---------------------------------------
        TC_PROF(t_p1 += clock64() - tp1;)
        for (int ph = 1; ph < 3; ++ph) {
          TC_PROF(const long long tph = clock64();)
          const int nb = ph == 2 ? a.n_pb : 1;
          for (int blk = 0; blk < nb; ++blk) {
            const int region = acc_it & 1;
            TC_PROF(const long long ta0 = clock64();)
            mbar_wait_cluster(&acc_empty[region], ((acc_it >> 1) & 1) ^ 1);
            TC_PROF(t_acc += clock64() - ta0;)
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + region * 256;
            for (int s = 0; s < KS; ++s, ++wit) {
              TC_PROF(const long long ti0 = clock64();)
              if (blk == 0) mbar_wait_cluster(&img_full[s], img_gen & 1);
              TC_PROF(if (ph == 1) t_i2 += clock64() - ti0; else t_i3 += clock64() - ti0;)
              slab(sbase + S::OFF_A + s * S::A_SLAB, wit % NWS, s == 0, d_tmem);
              if (lane == 0 && s == KS - 1) mma_commit_pair(&acc_full[region], cmask);
              __syncwarp();
            }
            ++acc_it;
          }
          ++img_gen;
          TC_PROF(if (ph == 1) t_p2 += clock64() - tph; else t_p3 += clock64() - tph;)
        }
        if (lane == 0) mma_commit_pair(img_free, cmask);
        __syncwarp();
      }
      TC_PROF(if (lane == 0 && blockIdx.x == 0) {
        g_tc16_stats[0] += (unsigned long long)(clock64() - t_begin); g_tc16_stats[1] += (unsigned long long)t_acc;
        g_tc16_stats[2] += (unsigned long long)t_a; g_tc16_stats[3] += (unsigned long long)t_w; g_tc16_stats[4] += 1ull;
        g_tc16_stats[5] += (unsigned long long)n_iter; g_tc16_stats[6] += (unsigned long long)t_i2;
        g_tc16_stats[7] += (unsigned long long)t_i3; g_tc16_stats[8] += (unsigned long long)t_p1;
        g_tc16_stats[9] += (unsigned long long)t_p2; g_tc16_stats[10] += (unsigned long long)t_p3;
      })
    }
  } else if (warp >= EPI_W) {
    // =========================== producers: [h | agg / div] -> fp16 hi | lo slabs ======================================
    const int pt = tid - EPI_T;
    const int chunk = pt & 7;
    const int rg = pt >> 3;
    uint32_t it = 0;
    for (int iter = 0; iter < n_iter; ++iter) {
      int row0, nrows;
      tile_of(iter, row0, nrows);
      const int rlast = nrows > 0 ? nrows - 1 : 0;
      const float* ph_[ROWS_PT];
      const float* pa_[ROWS_PT];
#pragma unroll
      for (int p = 0; p < ROWS_PT; ++p) {
        const size_t grow = (size_t)(row0 + min(rg * ROWS_PT + p, rlast));
        ph_[p] = a.h + grow * H + 4 * chunk;
        pa_[p] = a.agg + grow * H + 4 * chunk;
      }
      if (iter > 0) mbar_wait(img_free, (iter - 1) & 1);        // the ring aliases the image the previous tile's MMAs read
#pragma unroll 1
      for (int s = 0; s < 2 * KS; ++s, ++it) {
        const int st = it % NAS;
        const bool from_h = s < KS;
        const int k0 = (from_h ? s : s - KS) * BK;
        float4 v[ROWS_PT][2];
#pragma unroll
        for (int p = 0; p < ROWS_PT; ++p) {
          const float* src = (from_h ? ph_[p] : pa_[p]) + k0;
          v[p][0] = __ldg(reinterpret_cast<const float4*>(src));
          v[p][1] = __ldg(reinterpret_cast<const float4*>(src + KH));
        }
        mbar_wait(&a_empty[st], ((it / NAS) & 1) ^ 1);
        const uint32_t a_hi = sbase + S::OFF_A + st * S::A_SLAB;
#pragma unroll
        for (int p = 0; p < ROWS_PT; ++p) {
          float e[8] = {v[p][0].x, v[p][0].y, v[p][0].z, v[p][0].w, v[p][1].x, v[p][1].y, v[p][1].z, v[p][1].w};
          if (!from_h && a.agg_div != 1.0f) {
#pragma unroll
            for (int c = 0; c < 8; ++c) e[c] = __fdiv_rn(e[c], a.agg_div);
          }
          uint4 hi, lo;
          split_f16x8(e, hi, lo);
          const uint32_t off = sw128_off(rg * ROWS_PT + p, chunk);
          sts128(a_hi + off, hi);
          sts128(a_hi + TM * 128 + off, lo);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { if (crank != 0) mbar_arrive_remote(&a_full[st], 0); else mbar_arrive(&a_full[st]); }
      }
    }
  } else {
    // =========================== epilogue: thread = (row = TMEM lane, column half) ======================================
    const int r = (warp & 3) * 32 + lane;
    const int hf = warp >> 2;
    constexpr int HC = H / 2;
    constexpr int NCH = HC / 32;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + hf * HC;
    const float scale1 = __ldg(reinterpret_cast<const float*>(a.pack1)) * a.rz1;
    const float scale2 = __ldg(reinterpret_cast<const float*>(a.pack2)) * a.rz2;
    const float scale3 = __ldg(reinterpret_cast<const float*>(a.pack3)) * a.rz2;
    const int ld_pq = a.n_pb * H;
    uint32_t acc_it = 0;
    auto release_acc = [&](int region) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (crank != 0) mbar_arrive_remote(&acc_empty[region], 0); else mbar_arrive(&acc_empty[region]); }
    };
    // four consecutive values of this row (slab columns 4c..4c+3 of a 32-column piece) -> the fp16 hi | lo halves of the
    // operand image: the 16-byte chunk c of a slab row holds slab columns 4c..4c+3 (low 8 bytes) and 32+4c..32+4c+3 (high)
    auto put4 = [&](uint32_t img_piece, int c, f32x2 m0, f32x2 m1) {
      uint32_t h0, l0, h1, l1;
      split_f16x2(m0, h0, l0);
      split_f16x2(m1, h1, l1);
      const uint32_t off = img_piece + sw128_off(r, c);
      sts64(off, h0, h1);
      sts64(off + TM * 128, l0, l1);
    };
    auto piece_done = [&](int sl) {                // this warp's 32 rows x 32 columns of image slab sl are in place
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { if (crank != 0) mbar_arrive_remote(&img_full[sl], 0); else mbar_arrive(&img_full[sl]); }
    };
    for (int iter = 0; iter < n_iter; ++iter) {
      int row0, nrows;
      tile_of(iter, row0, nrows);
      // ---- epilogue 1: t1 = SiLU(scale D + b1) -> image ------------------------------------------------------------------
      {
        const int region = acc_it & 1;
        mbar_wait(&acc_full[region], (acc_it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tlane + region * 256;
#pragma unroll 1
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t v[32];
          tmem_ld32(taddr + cc * 32, v);
          tmem_ld_wait();
          if (cc == NCH - 1) release_acc(region);
          const int col0 = hf * HC + cc * 32;
          const uint32_t img_piece = sbase + S::OFF_A + (col0 / BK) * S::A_SLAB + ((col0 / 32) & 1) * 8;
          const f32x2 sc = pk2(scale1, scale1);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b1 + col0 + 4 * c4));
            f32x2 m0 = fma2(pk2(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1])), sc, pk2(b4.x, b4.y));
            f32x2 m1 = fma2(pk2(__uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3])), sc, pk2(b4.z, b4.w));
            silu_x4(m0, m1);
            put4(img_piece, c4, m0, m1);
          }
          piece_done(col0 / BK);
        }
        ++acc_it;
      }
      // Global rows move COALESCED: every 32 x 16 piece of the accumulator goes through this warp's transposition tile and is
      // then handled by lane (rl = row % 8, c4 = 16-byte column group): 8 rows x 64 contiguous bytes per instruction.
      const uint32_t Tw = sbase + S::OFF_T + warp * (32 * 20 * 4);
      const int rl = lane >> 2, c4l = lane & 3;
      const int wrow0 = row0 + (warp & 3) * 32;
      const int nval = min(32, nrows - (warp & 3) * 32);        // rows of this warp that exist (may be <= 0)
      auto to_tile = [&](const uint32_t (&v)[16]) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128f(Tw + (lane * 20 + 4 * q) * 4, make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                             __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])));
      };
      // ---- epilogue 2: h' = scale D + b2 + h -> global (fp32) and image; the consumed agg rows are handed back zeroed ----
      {
        const int region = acc_it & 1;
        mbar_wait(&acc_full[region], (acc_it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tlane + region * 256;
#pragma unroll 1
        for (int sc = 0; sc < HC / 16; ++sc) {
          uint32_t v[16];
          tmem_ld16(taddr + sc * 16, v);
          const int col = hf * HC + sc * 16 + 4 * c4l;            // this lane's 4 columns in the coalesced phase
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b2 + col));
          float4 rs[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int grow_c = min(wrow0 + max(0, min(8 * k + rl, nval - 1)), a.n_rows - 1);      // clamped: in bounds
            rs[k] = __ldg(reinterpret_cast<const float4*>(a.h + (size_t)grow_c * H + col));
          }
          tmem_ld_wait();
          if (sc == HC / 16 - 1) release_acc(region);
          to_tile(v);
          __syncwarp();
          const int lc = col % BK;                                 // column inside its k-slab
          const uint32_t img_piece = sbase + S::OFF_A + (col / BK) * S::A_SLAB + (lc / 32) * 8;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = 8 * k + rl;
            const float4 t = lds128f(Tw + (row * 20 + 4 * c4l) * 4);
            const float o0 = fmaf(t.x, scale2, b4.x) + rs[k].x, o1 = fmaf(t.y, scale2, b4.y) + rs[k].y;
            const float o2 = fmaf(t.z, scale2, b4.z) + rs[k].z, o3 = fmaf(t.w, scale2, b4.w) + rs[k].w;
            if (row < nval) {
              const size_t goff = (size_t)(wrow0 + row) * H + col;
              *reinterpret_cast<float4*>(a.h_out + goff) = make_float4(o0, o1, o2, o3);
              if (a.zero_buf) *reinterpret_cast<float4*>(a.zero_buf + goff) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t h0, l0, h1, l1;
            split_f16x2(pk2(o0, o1), h0, l0);
            split_f16x2(pk2(o2, o3), h1, l1);
            const uint32_t off = img_piece + sw128_off((warp & 3) * 32 + row, (lc % 32) / 4);
            sts64(off, h0, h1);
            sts64(off + TM * 128, l0, l1);
          }
          if (sc & 1) piece_done((hf * HC + sc * 16) / BK);          // 32 columns of image slab complete for this warp's rows
          else __syncwarp();
        }
        ++acc_it;
      }
      // ---- epilogue 3: pq block = scale D + b3 -> global ---------------------------------------------------------------------
      for (int blk = 0; blk < a.n_pb; ++blk) {
        const int region = acc_it & 1;
        mbar_wait(&acc_full[region], (acc_it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tlane + region * 256;
#pragma unroll 1
        for (int sc = 0; sc < HC / 16; ++sc) {
          uint32_t v[16];
          tmem_ld16(taddr + sc * 16, v);
          const int col = blk * H + hf * HC + sc * 16 + 4 * c4l;
          const float4 b4 = a.b3 ? __ldg(reinterpret_cast<const float4*>(a.b3 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          if (sc == HC / 16 - 1) release_acc(region);
          to_tile(v);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = 8 * k + rl;
            const float4 t = lds128f(Tw + (row * 20 + 4 * c4l) * 4);
            if (row < nval)
              *reinterpret_cast<float4*>(a.pq_out + (size_t)(wrow0 + row) * ld_pq + col) =
                  make_float4(fmaf(t.x, scale3, b4.x), fmaf(t.y, scale3, b4.y), fmaf(t.z, scale3, b4.z), fmaf(t.w, scale3, b4.w));
          }
          __syncwarp();
        }
        ++acc_it;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int H>
int launch_chain(const ChainArgs& a, cudaStream_t st) {
  using S = ChainSmem<H>;
  static DeviceOnce once;
  bool fresh;
  const int slot = device_slot(once, fresh);
  if (fresh) {
    cudaError_t e = cudaFuncSetAttribute(chain16_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::ALLOC);
    if (e != cudaSuccess) {
      set_error("chain16_kernel: cannot reserve %u bytes of shared memory: %s", S::ALLOC, cudaGetErrorString(e));
      return -2;
    }
    once.done[slot] = true;
  }
  const int sm_count = once.sm_count[slot];
  if (a.n_tile == 0) return 0;
  const int work = (a.n_tile + 1) / 2 * 2;
  int grid = work < sm_count ? work : sm_count;
  grid = grid / 2 * 2;
  ChainArgs args = a;
  args.rz1 = 1.0f + RZ_BIAS_PER_MMA * (float)(2 * S::KS * (BK / 16) * 3);
  args.rz2 = 1.0f + RZ_BIAS_PER_MMA * (float)(S::KS * (BK / 16) * 3);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CHAIN_THREADS);
  cfg.dynamicSmemBytes = S::ALLOC;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, chain16_kernel<H>, args);
  if (e != cudaSuccess) {
    set_error("chain16_kernel launch: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

// One warp per tile: {first receiver, first sender, staged, 0}.  Edge rows are sorted by (molecule, receiver, sender), so the
// first / last row hold the extreme receivers; senders are scanned.
__global__ void tile_meta_kernel(int n_tile, int tile_m, int n_edge, const int* __restrict__ tile_row,
                                 const int* __restrict__ edge_i, const int* __restrict__ edge_j, int4* __restrict__ out) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tile >= n_tile) return;
  const int r0 = tile_row ? tile_row[tile] : tile * tile_m;
  const int r1 = tile_row ? tile_row[tile + 1] : min(n_edge, r0 + tile_m);
  int jmin = INT_MAX, jmax = INT_MIN;
  for (int r = r0 + lane; r < r1; r += 32) { const int j = edge_j[r]; jmin = min(jmin, j); jmax = max(jmax, j); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
    jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
  }
  if (lane == 0) {
    int4 m = make_int4(0, 0, 0, 0);
    if (r1 > r0) {
      const int ilo = edge_i[r0], ihi = edge_i[r1 - 1];
      m.x = ilo; m.y = jmin;
      m.z = (ihi - ilo < (int)STAGE_PBOX && jmax - jmin < (int)STAGE_QBOX && ihi >= ilo) ? 1 : 0;
    }
    out[tile] = m;
  }
}
}  // namespace

int launch_edge_tc16(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                     const float* pq, int pq_ld, const float* x, const float* x0, const float* r_edge,
                     const float* d0_edge, const float* u_edge, float* out, cudaStream_t st) {
  GEOLDM_REQUIRE(b.tile_m == TM, "edge_tc16: batch tile_m=%d, kernel needs %d", b.tile_m, TM);
  GEOLDM_REQUIRE(w.tc_pack != nullptr, "edge_tc16: tc_pack missing (weights not packed for the tensor-core path)");
  GEOLDM_REQUIRE(equiv || !cfg.attention || w.b_out != nullptr, "edge_tc16: attention needs b_out");
  GEOLDM_REQUIRE(cfg.hidden_nf % BK == 0, "edge_tc16: hidden_nf %d must be a multiple of %d", cfg.hidden_nf, BK);
  Args a{};
  a.n_tile = b.n_tile; a.n_rows = b.n_edge; a.tile_row = b.tile_row; a.n_blocks = 1;
  a.n_slabs = cfg.hidden_nf / BK;
  a.pq = pq; a.pq_ld = pq_ld; a.x = x; a.x0 = x0; a.edge_i = b.edge_i; a.edge_j = b.edge_j; a.w_rd = w.w_rd;
  a.r_edge = (r_edge && d0_edge) ? r_edge : nullptr; a.d0_edge = d0_edge;
  a.u_edge = reinterpret_cast<const float4*>(u_edge);
  a.w_pack = reinterpret_cast<const uint8_t*>(w.tc_pack);
  a.b2 = w.b2; a.w_out = w.w_out; a.b_out = w.b_out; a.out = out;
  a.norm_constant = cfg.norm_constant; a.coords_range = cfg.coords_range;
  a.attention = cfg.attention; a.use_tanh = cfg.tanh;
  // per-tile staging of the projection rows through TMA when the batch carries the table (GEOLDM_TC_STAGE=0: gather path)
  static int stage = -1;
  if (stage < 0) { const char* e = getenv("GEOLDM_TC_STAGE"); stage = e ? atoi(e) : 1; }
  StageMaps sm;
  const StageMaps* smp = nullptr;
  if (stage && b.tile_meta != nullptr && b.n_tile > 0 && (reinterpret_cast<uintptr_t>(pq) & 15) == 0 && pq_ld % 4 == 0) {
    if (int rc = make_stage_maps(pq, pq_ld, b.n_node, cfg.hidden_nf, sm)) return rc;
    a.tile_meta = reinterpret_cast<const int4*>(b.tile_meta);
    smp = &sm;
  }
  return equiv ? launch_h<MODE_EQUIV>(cfg.hidden_nf, a, st, smp) : launch_h<MODE_GCL>(cfg.hidden_nf, a, st, smp);
}

int launch_linear_tc16(int H, const float* a1, int k1, const float* a2, int k2, float a2_div, const void* w_pack,
                       int n_blocks, const float* bias, const float* res, int epi, float* out, int m, cudaStream_t st,
                       float* zero_buf) {
  GEOLDM_REQUIRE(zero_buf == nullptr || epi == 2, "linear_tc16: zero_buf needs the residual epilogue (epi 2)");
  GEOLDM_REQUIRE(k1 % BK == 0 && k2 % BK == 0 && k1 + k2 > 0, "linear_tc16: k1=%d k2=%d must be multiples of %d", k1, k2, BK);
  GEOLDM_REQUIRE(w_pack != nullptr, "linear_tc16: w_pack missing");
  Args a{};
  a.n_tile = (m + TM - 1) / TM; a.n_rows = m; a.tile_row = nullptr; a.n_blocks = n_blocks;
  a.n_slabs = (k1 + k2) / BK;
  a.a1 = a1; a.a2 = a2; a.k1 = k1; a.k2 = k2; a.a2_div = a2_div;
  a.w_pack = reinterpret_cast<const uint8_t*>(w_pack);
  a.b2 = bias; a.res = res; a.epi = epi; a.out = out; a.ldo = n_blocks * H; a.zero_buf = zero_buf;
  return launch_h<MODE_DENSE>(H, a, st);
}

// out[m][H] = a[m][H] (row stride ld) * W^T with a power-of-two pre-scaling of A taken from its own maximum: the
// input-gradient GEMM dX = dY W of a square Linear layer (dY is 1 / batch-size small: unscaled it leaves the normal fp16
// range and the hi | lo split loses its low half).  amax_scratch: 4 bytes of device memory.
int launch_linear_tc16_grad(int H, const float* a_rows, int ld, const void* w_pack, float* out, int m, unsigned* amax_scratch,
                            int amax_ready, cudaStream_t st) {
  GEOLDM_REQUIRE(w_pack != nullptr && amax_scratch != nullptr, "linear_tc16_grad: w_pack / amax_scratch missing%s", "");
  GEOLDM_REQUIRE(ld % 4 == 0 && ld >= H && (reinterpret_cast<uintptr_t>(a_rows) & 15) == 0, "linear_tc16_grad: ld=%d", ld);
  if (m == 0) return 0;
  GEOLDM_REQUIRE(ld == H, "linear_tc16_grad: the operand must be dense (ld=%d, H=%d)", ld, H);
  if (!amax_ready) {
    cudaError_t e = cudaMemsetAsync(amax_scratch, 0, sizeof(unsigned), st);
    GEOLDM_REQUIRE(e == cudaSuccess, "linear_tc16_grad: memset: %s", cudaGetErrorString(e));
    const size_t tot = (size_t)m * H;
    pack16_amax_kernel<<<(unsigned)((tot + 1023) / 1024 < 592 ? (tot + 1023) / 1024 : 592), 256, 0, st>>>(a_rows, tot, amax_scratch);
    GEOLDM_CHECK_LAUNCH("pack16_amax_kernel(grad)");
  }
  Args a{};
  a.n_tile = (m + TM - 1) / TM; a.n_rows = m; a.tile_row = nullptr; a.n_blocks = 1; a.n_slabs = H / BK;
  a.pq = a_rows; a.pq_ld = ld; a.edge_i = nullptr; a.w_pack = reinterpret_cast<const uint8_t*>(w_pack); a.out = out; a.ldo = H;
  a.a_amax_bits = amax_scratch;
  return launch_h<MODE_RAW>(H, a, st);
}

int launch_node_chain16(int H, const float* h, const float* agg, float agg_div, const void* pack1, const float* b1,
                        const void* pack2, const float* b2, const void* pack3, const float* b3, int n_pb, float* h_out,
                        float* pq_out, float* zero_buf, int m, cudaStream_t st) {
  GEOLDM_REQUIRE(pack1 && pack2 && pack3 && b1 && b2, "node_chain16: weight packs / biases missing");
  GEOLDM_REQUIRE(n_pb >= 1 && n_pb <= 8, "node_chain16: %d projection blocks", n_pb);
  GEOLDM_REQUIRE(h != h_out, "node_chain16: h_out must not alias h (rows are re-read as the residual)");
  ChainArgs a{};
  a.n_tile = (m + TM - 1) / TM; a.n_rows = m;
  a.h = h; a.agg = agg; a.agg_div = agg_div;
  a.pack1 = reinterpret_cast<const uint8_t*>(pack1); a.b1 = b1;
  a.pack2 = reinterpret_cast<const uint8_t*>(pack2); a.b2 = b2;
  a.pack3 = reinterpret_cast<const uint8_t*>(pack3); a.b3 = b3; a.n_pb = n_pb;
  a.h_out = h_out; a.pq_out = pq_out; a.zero_buf = zero_buf;
  switch (H) {
    case 64: return launch_chain<64>(a, st);
    case 128: return launch_chain<128>(a, st);
    case 192: return launch_chain<192>(a, st);
    case 256: return launch_chain<256>(a, st);
    default: set_error("node_chain16 supports hidden_nf 64/128/192/256, got %d", H); return -1;
  }
}

int launch_tc16_selftest(int H, const float* pq, const int* edge_i, const int* tile_row, int n_tile, int n_rows,
                         const void* w_pack, float* out, cudaStream_t st) {
  Args a{};
  a.n_tile = n_tile; a.n_rows = n_rows; a.tile_row = tile_row; a.n_blocks = 1; a.n_slabs = H / BK;
  a.pq = pq; a.pq_ld = 2 * H; a.edge_i = edge_i; a.w_pack = reinterpret_cast<const uint8_t*>(w_pack); a.out = out; a.ldo = H;
  return launch_h<MODE_RAW>(H, a, st);
}

}  // namespace geoldm

extern "C" {
int geoldm_batch_tile_meta(const geoldm_batch* b, int* out, void* stream) {
  GEOLDM_REQUIRE(b != nullptr && out != nullptr, "batch_tile_meta: null argument");
  if (b->n_tile == 0) return 0;
  geoldm::tile_meta_kernel<<<(b->n_tile + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
      b->n_tile, b->tile_m, b->n_edge, b->tile_row, b->edge_i, b->edge_j, reinterpret_cast<int4*>(out));
  GEOLDM_CHECK_LAUNCH("tile_meta_kernel");
  return 0;
}
/* debug (GEOLDM_TC_PROFILE builds): read + reset the 16 cycle counters of the fp16-split kernel; synchronises */
int geoldm_tc16_read_stats(unsigned long long* host_out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_out, geoldm::g_tc16_stats, sizeof(unsigned long long) * 16);
  unsigned long long zero[16] = {0};
  cudaMemcpyToSymbol(geoldm::g_tc16_stats, zero, sizeof(zero));
  return 0;
}
size_t geoldm_tc_pack16_bytes(int H, int n_out, int k) {
  (void)H;
  return geoldm::PACK_HDR + (size_t)n_out * k * 2 * sizeof(__half);
}

static int tc_pack16_impl(int H, const float* w, int n_out, int k, void* pack, void* stream, int transposed);
int geoldm_tc_pack16(int H, const float* w, int n_out, int k, void* pack, void* stream) {
  return tc_pack16_impl(H, w, n_out, k, pack, stream, 0);
}
int geoldm_tc_pack16_t(int H, const float* w_kn, int n_out, int k, void* pack, void* stream) {
  return tc_pack16_impl(H, w_kn, n_out, k, pack, stream, 1);
}
int geoldm_tc_pack16_pair(int H, const float* w, void* pack_fwd, void* pack_t, void* stream) {
  GEOLDM_REQUIRE(H % 64 == 0 && H <= 256 && H > 0, "tc_pack16_pair: H=%d", H);
  GEOLDM_REQUIRE(w != nullptr && (pack_fwd != nullptr || pack_t != nullptr), "tc_pack16_pair: null argument%s", "");
  geoldm::pack16_pair_kernel<<<H * H / geoldm::PAIR_THREADS, geoldm::PAIR_THREADS, 0, (cudaStream_t)stream>>>(H, w, reinterpret_cast<uint8_t*>(pack_fwd),
                                                                   reinterpret_cast<uint8_t*>(pack_t));
  GEOLDM_CHECK_LAUNCH("pack16_pair_kernel");
  return 0;
}
int geoldm_linear_tc_grad(int H, const float* dy, int ld, const void* w_pack, float* out, int m, void* amax_scratch,
                          int amax_ready, void* stream) {
  return geoldm::launch_linear_tc16_grad(H, dy, ld, w_pack, out, m, reinterpret_cast<unsigned*>(amax_scratch), amax_ready,
                                         (cudaStream_t)stream);
}
static int tc_pack16_impl(int H, const float* w, int n_out, int k, void* pack, void* stream, int transposed) {
  GEOLDM_REQUIRE(H % 64 == 0 && H <= 256 && n_out % H == 0 && k % 64 == 0, "tc_pack16: H=%d n_out=%d k=%d", H, n_out, k);
  const size_t tot = (size_t)n_out * k;
  if (tot == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(pack, 0, geoldm::PACK_HDR, st);
  if (e != cudaSuccess) { geoldm::set_error("tc_pack16: memset: %s", cudaGetErrorString(e)); return -2; }
  unsigned* amax = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(pack) + 4);
  geoldm::pack16_amax_kernel<<<(unsigned)((tot + 1023) / 1024 < 592 ? (tot + 1023) / 1024 : 592), 256, 0, st>>>(w, tot, amax);
  GEOLDM_CHECK_LAUNCH("pack16_amax_kernel");
  geoldm::pack16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(H, w, n_out, k, reinterpret_cast<uint8_t*>(pack), transposed);
  GEOLDM_CHECK_LAUNCH("pack16_kernel");
  return 0;
}
}

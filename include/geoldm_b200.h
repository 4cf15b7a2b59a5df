/*
 * geoldm_b200.h — C ABI of the B200-native GeoLDM EGNN-denoiser sampling path.
 *
 * The reference (mint258/GeoLDM) is pure Python/PyTorch and has NO plugin/FFI layer for this path
 * (SURVEY.md §8b): the boundary there is the Python call
 *     EnVariationalDiffusion.phi -> self.dynamics._forward(t, xh, node_mask, edge_mask, context)
 *                                                   (equivariant_diffusion/en_diffusion.py:314-317)
 * and, one level up, generative_model.sample(...)   (qm9/sampling.py:139).
 * This header is therefore the interface a maintainer binds (ctypes stub in INTEGRATION.md) to
 * replace the bodies of the functions cited on each entry point below.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - all floating point is fp32, indices int32, row-major, ragged-packed (no padding): real
 *     nodes of all molecules are concatenated, N = sum n_b; edges are implicit (fully connected
 *     within a molecule, i != j) and described by the geoldm_batch tables;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*), performs no
 *     allocation and no host synchronisation (CUDA-graph capturable); workspace is caller-owned;
 *   - return value 0 = ok, <0 = error; geoldm_last_error() returns the text (thread-local).
 */
#ifndef GEOLDM_B200_H
#define GEOLDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEOLDM_MAX_LAYERS 16
#define GEOLDM_MAX_SUBLAYERS 4
#define GEOLDM_ABI_VERSION 4

/* arithmetic mode of the 256x256 edge/node contractions */
enum {
  GEOLDM_MMA_FP32_SIMT = 0,   /* fp32 FFMA on CUDA cores (exact fp32 products)               */
  GEOLDM_MMA_3XTF32 = 1,      /* tcgen05 kind::tf32, hi*hi + hi*lo + lo*hi, fp32 accumulate  */
  GEOLDM_MMA_TF32 = 2,        /* tcgen05 kind::tf32 single pass (fast mode, fails 1e-5 gate) */
  GEOLDM_MMA_BF16 = 3,        /* tcgen05 kind::f16 bf16 single pass (not implemented)        */
  GEOLDM_MMA_3XF16 = 4        /* tcgen05 kind::f16, fp16 hi/lo split of both operands (3 products, fp32 accumulate):
                                 same 22-bit products as 3xTF32 at twice the MMA rate; needs |activation| < 65504;
                                 weight images from geoldm_tc_pack16                          */
};

/* Mirrors the ctor kwargs of EGNN / EGNN_dynamics_QM9 (egnn/egnn_new.py:151-153, egnn/models.py:9-13). */
typedef struct {
  int hidden_nf;             /* nf; one of 32, 64, 128, 192, 256                            */
  int n_layers;              /* number of EquivariantBlocks                                 */
  int inv_sublayers;         /* GCLs per block                                              */
  int in_node_nf;            /* features entering `embedding` (latent + time + context)     */
  int out_node_nf;           /* features leaving `embedding_out`                            */
  int attention;             /* GCL attention gate                                          */
  int tanh;                  /* tanh-bounded coordinate message                             */
  float norm_constant;       /* coord2diff denominator constant                             */
  float coords_range;        /* 15.0 per block (not /n_layers; egnn_new.py:160 vs :178)     */
  float agg_div;             /* 'sum': normalization_factor; 'mean': padded n_max           */
  int mma_mode;              /* GEOLDM_MMA_*                                                */
} geoldm_egnn_config;

/* One edge MLP (GCL.edge_mlp+att_mlp, egnn_new.py:14-28; or EquivariantUpdate.coord_mlp, :75-82),
 * with the first layer split per node:  W1 [H,2H+2] -> pq_wt = [W1[:, :H]^T | W1[:, H:2H]^T]. */
typedef struct {
  const float* pq_wt;        /* [H][2H]  (k-major: row k holds the 2H outputs)              */
  const float* pq_b;         /* [2H]     first-layer bias in [0,H), zeros in [H,2H)         */
  const float* w_rd;         /* [2][H]   W1[:, 2H] (current d^2), W1[:, 2H+1] (entry d^2)   */
  const float* w2t;          /* [H][H]   second layer, transposed (k-major)                 */
  const float* b2;           /* [H]                                                         */
  const float* w_out;        /* [H]      att_mlp.0.weight  |  coord_mlp.4.weight            */
  const float* b_out;        /* [1]      att_mlp.0.bias    |  NULL                          */
  const void* tc_pack;       /* tensor-core pack of the second layer  (geoldm_tc_pack_bytes(H, H, H))   */
  const void* tc_pack_pq;    /* tensor-core pack of the split first layer, 2 column blocks of H          */
} geoldm_edge_mlp;

typedef struct {
  geoldm_edge_mlp edge;
  const float* node_w1t;     /* [2H][H]  node_mlp.0 transposed: rows [0,H) act on h, [H,2H) on agg */
  const float* node_b1;      /* [H]                                                         */
  const float* node_w2t;     /* [H][H]   node_mlp.2 transposed                              */
  const float* node_b2;      /* [H]                                                         */
  const void* tc_pack_node1; /* tensor-core pack of node_mlp.0 (K = 2H)                     */
  const void* tc_pack_node2; /* tensor-core pack of node_mlp.2 (K = H)                      */
} geoldm_gcl;

typedef struct {
  geoldm_gcl gcl[GEOLDM_MAX_SUBLAYERS];
  geoldm_edge_mlp equiv;
  /* tensor-core path: this block's coord_mlp P|Q and the NEXT block's gcl_0 P|Q read the same h, so they are one
   * GEMM with 4 column blocks: rows [equiv src; equiv dst; next gcl_0 src; next gcl_0 dst] (NULL in the last block) */
  const void* tc_pack_pq4;
  const float* pq4_b;        /* [4H] = [equiv b1 | 0 | next gcl_0 b1 | 0] */
} geoldm_block;

typedef struct {
  const float* emb_w;        /* [H][in_node_nf]  (PyTorch layout)                           */
  const float* emb_b;        /* [H]                                                         */
  const float* out_w;        /* [out_node_nf][H]                                            */
  const float* out_b;        /* [out_node_nf]                                               */
  geoldm_block block[GEOLDM_MAX_LAYERS];
} geoldm_egnn_weights;

/* Ragged batch tables (replace get_adj_matrix + node_mask/edge_mask, egnn/models.py:115-134,
 * qm9/sampling.py:118-128).  Edge rows are sorted by (molecule, receiver i, sender j != i). */
typedef struct {
  int n_mol, n_node, n_edge, n_tile, tile_m;
  const int* mol_off;        /* [n_mol+1] first node of each molecule                       */
  const int* node_mol;       /* [n_node]  molecule of each node                             */
  const int* edge_i;         /* [n_edge]  receiver (aggregation index, `row`)               */
  const int* edge_j;         /* [n_edge]  sender (`col`)                                    */
  const int* tile_row;       /* [n_tile+1] first edge row of each tile; rows per tile <= tile_m */
  const int* tile_meta;      /* [n_tile][4] {first receiver, first sender, staged flag, 0} written by
                              * geoldm_batch_tile_meta (ABI v4), or NULL: the tcgen05 edge kernels then gather the
                              * projection rows per edge instead of staging them per tile through TMA          */
} geoldm_batch;

/* Per-tile staging table of the tcgen05 edge kernels (ABI v4).  For every tile of b->tile_m edge rows: the first
 * receiver and the first sender node it touches, and whether all its receivers / senders lie inside the boxes
 * (16 / 64 consecutive nodes) that one TMA tensor copy per k-slab brings into shared memory.  b->tile_meta is ignored
 * on input; `out` ([n_tile][4] int32, device) is what the caller stores there afterwards. */
int geoldm_batch_tile_meta(const geoldm_batch* b, int* out, void* stream);

int geoldm_abi_version(void);
const char* geoldm_last_error(void);
/* 1 if the library was built with the tcgen05 kernels for sm_100a */
int geoldm_has_tcgen05(void);

/* ---- whole-network entry point: EGNN.forward (egnn/egnn_new.py:184-197) ------------------- */
/* bytes of caller-owned scratch for a batch of n_node atoms and n_edge directed edges: node buffers (h, t1, agg, the
 * [N][4H] projections, coordinates) plus two [E] arrays of squared distances (ABI v2; v1 took n_node only) */
size_t geoldm_egnn_workspace_bytes(const geoldm_egnn_config* cfg, int n_node, int n_edge);
int geoldm_egnn_forward(const geoldm_egnn_config* cfg, const geoldm_egnn_weights* w, const geoldm_batch* b,
                        const float* h_in,  /* [N][in_node_nf] */
                        const float* x_in,  /* [N][3] */
                        float* h_out,       /* [N][out_node_nf] */
                        float* x_out,       /* [N][3] final coordinates x_in + dx */
                        float* dx_out,      /* [N][3] accumulated displacement sum_b agg_b (may be NULL) */
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- EGNN_dynamics_QM9._forward glue (egnn/models.py:56-76 and :80-113) ------------------- */
/* Gathers real nodes from the caller's (possibly padded) xh, appends time/context features.
 * node_src[k] = flat row of xh holding ragged node k.  t_mol: per-molecule time [n_mol] or NULL;
 * if NULL the time is t_table[*step_idx_dev] (sampler) .  context: [rows][ctx_nf] or NULL. */
int geoldm_dynamics_prep(const geoldm_batch* b, const int* node_src, const float* xh, int xh_dim,
                         const float* t_mol, const float* t_table, const int* step_idx_dev,
                         const float* context, int ctx_nf, int condition_time,
                         float* h_in, int in_node_nf, float* x, void* stream);
/* vel [N][3] = dx_out of geoldm_egnn_forward (dynamics: x_final - x, models.py:80) or x_out (decoder, :356).
 * finish_a: NaN scan over the whole batch (models.py:100-102), nan_flag must be zero on entry.
 * finish_b: zeroes vel if the flag is set, remove_mean_with_mask (utils.py:31-38), drops the time/context
 * columns (keeps h_keep), writes out[node_src[k]] (node_src NULL: ragged output). */
int geoldm_dynamics_finish_a(const geoldm_batch* b, const float* vel, int* nan_flag, void* stream);
int geoldm_dynamics_finish_b(const geoldm_batch* b, const int* node_src, const float* vel, const float* h_out,
                             int h_stride, int h_keep, const int* nan_flag, float* out, int out_dim,
                             void* stream);

/* ---- sampler (equivariant_diffusion/en_diffusion.py:716-795, 1099-1122) -------------------- */
/* z_s = z_t/alpha_ts - c_eps*eps_hat + sigma*noise, then CoM projection of the x part.
 * coef: [n_steps][4] = {alpha_ts, c_eps, sigma, t}; row *step_idx_dev is used.  noise is either
 * caller-provided (injected, ragged layout of z; draw d lives at noise + d*noise_stride floats, d =
 * *draw_idx_dev) or NULL -> Philox4x32-10 keyed by (seed, mol_id[m]),
 * counter (draw index *draw_idx_dev, node, column); noise x-part is mean-centred per molecule.
 * mode 0: p(z_s|z_t) update;  mode 1: p(x|z_0): out = (z - sigma0*eps)/alpha0 + sigma_x*noise with
 * coef row = {alpha0, sigma0, sigma_x, 0};  mode 2: out = noise (initial z_T). */
int geoldm_sampler_update(const geoldm_batch* b, int mode, const float* coef, const int* step_idx_dev,
                          const float* z, const float* eps_hat, const float* noise, size_t noise_stride, int dim,
                          uint64_t seed, const int64_t* mol_id, const int* draw_idx_dev,
                          float* z_out, void* stream);
/* *step_idx_dev += d_step; *draw_idx_dev += d_draw (single thread; keeps the loop graph-replayable) */
int geoldm_sampler_advance(int* step_idx_dev, int d_step, int* draw_idx_dev, int d_draw, void* stream);

/* ---- unit-level entry points (used by tests; same kernels as the forward) ------------------ */
/* a1+a3 edge part: agg[i] += sum_j e_ij (raw sum; caller zeroes agg and divides by agg_div) */
int geoldm_edge_gcl(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b,
                    const float* pq, const float* x, const float* x0, float* agg, void* stream);
/* a1 (coord2diff, egnn/egnn_new.py:249-255) for every packed edge row e = (i, j): r_out[e] = |x_i - x_j|^2 and, when
 * u_out != NULL, u_out[4e .. 4e+2] = (x_i - x_j) / (sqrt(r + 1e-8) + norm_constant), u_out[4e+3] = 0.  geoldm_egnn_forward
 * runs this once per block (mma_mode 3xf16) and feeds the results to the fused edge kernels below. */
int geoldm_edge_dist(const geoldm_batch* b, const float* x, float* r_out, float* u_out, float norm_constant,
                     void* stream);
/* the fused edge kernels exactly as geoldm_egnn_forward launches them in mma_mode 3xf16: per-edge squared distances of
 * the current (r_edge) and the EGNN-entry (d0_edge) coordinates and, for the coordinate update, the normalised differences
 * u_edge ([E][4]) are read instead of being recomputed from x; pq is [N][pq_ld] (P in columns [0,H), Q in [H,2H)). */
int geoldm_edge_gcl_pre(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b, const float* pq,
                        int pq_ld, const float* r_edge, const float* d0_edge, float* agg, void* stream);
int geoldm_edge_equiv_pre(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b,
                          const float* pq, int pq_ld, const float* r_edge, const float* d0_edge, const float* u_edge,
                          float* xagg, void* stream);
/* a1+a4: xagg[i] += sum_j u_ij * tanh(s_ij) * coords_range (raw sum) */
int geoldm_edge_equiv(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b,
                      const float* pq, const float* x, const float* x0, float* xagg, void* stream);
/* out[M][N] = epi(a1[M][k1] * wt[0:k1] + (a2[M][k2]/a2_div) * wt[k1:k1+k2] + bias); epi: 0 none, 1 SiLU,
 * 2 residual (out = res + ...).  wt is k-major [k1+k2][N]. */
int geoldm_linear(const float* a1, int k1, const float* a2, int k2, float a2_div, const float* wt,
                  const float* bias, const float* res, int epi, float* out, int m, int n, int mma_mode,
                  void* stream);
/* Tensor-core operand pack of a weight matrix W [n_out][k] (row-major, PyTorch Linear layout) for the tcgen05
 * kernels: n_out/H column blocks x k/32 k-slabs x {tf32-hi, tf32-lo} images of H rows x 128 bytes in the canonical
 * SWIZZLE_128B K-major order (the exact bytes the TMA engine drops into shared memory).  Runs on `stream`. */
size_t geoldm_tc_pack_bytes(int H, int n_out, int k);
int geoldm_tc_pack(int H, const float* w, int n_out, int k, void* pack, void* stream);
/* fp16-split pack for GEOLDM_MMA_3XF16: 128-byte header {float 2^-e, ...} followed by n_out/H column blocks x k/64
 * k-slabs x 2 N-halves x {fp16-hi, fp16-lo} images of W * 2^e (e chosen on the device so that max|W| 2^e is in
 * [2^13, 2^14)), same SWIZZLE_128B K-major order.  H and k multiples of 64. */
size_t geoldm_tc_pack16_bytes(int H, int n_out, int k);
int geoldm_tc_pack16(int H, const float* w, int n_out, int k, void* pack, void* stream);
/* Fused node chain of one GCL (egnn/egnn_new.py:47-56) plus the first-layer projections reading its output, fp16-split
 * tcgen05 arithmetic, ONE persistent launch (ABI v4):
 *   t1 = SiLU([h | agg / agg_div] W1^T + b1);  h_out = h + t1 W2^T + b2;  pq_out[:, blk H ..] = h_out Wp_blk^T + b3
 * pack1 / pack2 / pack3: geoldm_tc_pack16 images of node_mlp.0 (n_out H, k 2H), node_mlp.2 (H, H) and of n_blocks3 stacked
 * [H][H] projection blocks; pq_out is [m][n_blocks3 H]; agg_zero (may be NULL) = an [m][H] buffer to clear (the consumed
 * agg); h_out must not alias h.  H in {64,128,192,256}. */
int geoldm_node_chain(int H, const float* h, const float* agg, float agg_div, const void* pack1, const float* b1,
                      const void* pack2, const float* b2, const void* pack3, const float* b3, int n_blocks3, float* h_out,
                      float* pq_out, float* agg_zero, int m, void* stream);

/* same contract as geoldm_linear with n = n_blocks*H outputs, on the tensor cores; terms: 3 = 3xTF32, 1 = TF32,
 * 16 = 3xF16 (w_pack from geoldm_tc_pack16) */
int geoldm_linear_tc(int H, int terms, const float* a1, int k1, const float* a2, int k2, float a2_div,
                     const void* w_pack, int n_blocks, const float* bias, const float* res, int epi, float* out,
                     int m, void* stream);
/* input-gradient GEMM of a square Linear layer on the fp16-split tensor-core kernel:  out[m][H] = dy[m][H] * W  with W
 * [H out-features][H in-features] packed by geoldm_tc_pack16_t (the transposed operand).  dy is a GRADIENT (1 / batch-size
 * small): the kernel scales it by the power of two that puts max|dy| in [2^13, 2^14) before the fp16 hi | lo split and
 * divides it out in the epilogue (both exact), so the result is as accurate as the forward GEMM regardless of the
 * magnitude of dy.  ld == H; amax_scratch: 4 bytes of device memory, overwritten with the bit pattern of max|dy| unless
 * amax_ready != 0: then it already holds it (written by the kernel that produced dy, geoldm_train_edge_tail_bwd). */
int geoldm_tc_pack16_t(int H, const float* w_kn, int n_out, int k, void* pack, void* stream);
/* both operand images of a SQUARE weight W [H][H] in one single-block launch: pack_fwd as geoldm_tc_pack16(H, W, H, H),
 * pack_t as geoldm_tc_pack16_t(H, W, H, H) (either may be NULL); for weights that are re-packed every optimiser step */
int geoldm_tc_pack16_pair(int H, const float* w, void* pack_fwd, void* pack_t, void* stream);
int geoldm_linear_tc_grad(int H, const float* dy, int ld, const void* w_pack, float* out, int m, void* amax_scratch,
                          int amax_ready, void* stream);
/* descriptor / swizzle / pipeline self-test: out[row][0:H] = sum_k a[src_row[row]][k] * W[:, k] with a row stride
 * of 2H floats (the P|Q layout), rows split into tiles by tile_row like the edge kernels */
int geoldm_tc_selftest(int H, int terms, const float* a, const int* src_row, const int* tile_row, int n_tile,
                       int n_rows, const void* w_pack, float* out, void* stream);
/* ---- training side: weight-gradient GEMM  C[n][k] += A[m][n]^T * B[m][k]  (fp32 FFMA, caller
 * zeroes C; lda/ldb multiples of 4).  Forward and input-gradient GEMMs of a Linear layer use geoldm_linear. */
int geoldm_gemm_tn(const float* a, int lda, const float* b, int ldb, float* c, int ldc, int m, int n, int k,
                   void* stream);
/* the same GEMM with the bias gradient of the layer on the side: colsum[n] += sum_m A[m][n] (caller zeroes colsum; NULL:
 * plain geoldm_gemm_tn).  One pass over dY instead of a separate column-sum launch per Linear layer. */
int geoldm_gemm_tn_bias(const float* a, int lda, const float* b, int ldb, float* c, int ldc, float* colsum, int m, int n,
                        int k, void* stream);
/* Fused element-wise stages of the edge MLPs for the autograd path (forward + backward, everything recomputed from the
 * GEMM operands; one warp per edge; H <= 256, H a multiple of 4, [.][H] operands 16-byte aligned).  pq [N][pq_ld] = P | Q projections; r, d0 [E] squared distances;
 * w_rd [2][H]; outputs of the backward kernels that are sums over edges (dpq, dw_rd, db2, dw, dbw) are ACCUMULATED with
 * atomics into caller-zeroed buffers.
 *   act : a[e] = SiLU(P[i_e] + Q[j_e] + r_e w_r + d0_e w_d)
 *   tail: m = SiLU(mpre + b2); gate != 0: agg[i_e] += m * (attention ? sigmoid(m.w + bw) : 1) / div;
 *                              gate == 0: sc[e] = m.w                       (egnn_new.py:30-45, 86-90, 258-267) */
int geoldm_train_edge_act_fwd(int n_edge, int H, const float* pq, int pq_ld, const float* r, const float* d0,
                              const float* w_rd, const int* edge_i, const int* edge_j, float* a, void* stream);
int geoldm_train_edge_act_bwd(int n_edge, int H, const float* pq, int pq_ld, const float* r, const float* d0,
                              const float* w_rd, const int* edge_i, const int* edge_j, const float* da, float* dpq,
                              float* dr, float* dd0, float* dw_rd, void* stream);
int geoldm_train_edge_tail_fwd(int n_edge, int H, const float* mpre, const float* b2, const float* w, const float* bw,
                               int gate, int attention, const int* edge_i, float div, float* agg, float* sc,
                               void* stream);
int geoldm_train_edge_tail_bwd(int n_edge, int H, const float* mpre, const float* b2, const float* w, const float* bw,
                               int gate, int attention, const int* edge_i, float div, const float* dagg, const float* dsc,
                               float* dmpre, float* db2, float* dw, float* dbw,
                               double* dbw_scratch, /* [geoldm_train_bwd_blocks(n_edge) + 1] doubles, ZEROED once (the kernel leaves it zeroed), or
                                                     * NULL: dbw is accumulated with float atomics in arrival order (not reproducible) */
                               void* damax,         /* 4 bytes, caller-zeroed, or NULL: bit pattern of max |dmpre| (what
                                                     * geoldm_linear_tc_grad(..., amax_ready = 1) scales its operand by) */
                               void* stream);
/* AdamW(amsgrad) + EMA of the weights as ONE multi-tensor launch (qm9/models.py:169-175, utils.py:5-28,
 * train_test.py:60-66).  table: one entry per parameter tensor (device array); chunk_map: n_chunks pairs {tensor index,
 * first element} covering every tensor in pieces of geoldm_optim_chunk() elements (device array); step: device scalar, the
 * update count INCLUDING this update (also stored to every entry's step); grad_scale: device scalar multiplied into the
 * gradients (the clipping coefficient) or NULL; ema / vmax / step of an entry may be NULL.
 *   p -= lr wd p;  m += (1 - b1)(g - m);  v = b2 v + (1 - b2) g^2;  vmax = max(vmax, v);
 *   p -= lr / (1 - b1^step) * m / (sqrt(vmax) / sqrt(1 - b2^step) + eps);  ema = ema beta + (1 - beta) p */
typedef struct geoldm_optim_tensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* vmax;
  float* ema;
  float* step;
  int n;
  int pad_;
} geoldm_optim_tensor;
int geoldm_adamw_ema_step(const geoldm_optim_tensor* table, const int* chunk_map, int n_chunks, const float* step,
                          const float* grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                          int amsgrad, float ema_beta, void* stream);
int geoldm_optim_chunk(void);
/* coord2diff and the coordinate update of EquivariantUpdate for the autograd path, forward and backward (one thread per
 * edge; egnn_new.py:249-255, 91-99):
 *   r[e] = |x_i - x_j|^2, u[e][3] = (x_i - x_j) / (sqrt(r + 1e-8) + norm_constant)   (u may be NULL: r only)
 *   bwd: gx [N][3] (caller-zeroed) += the gradient through r (gr [E], or NULL) and u (gu [E][3], or NULL)
 *   step[i_e][3] (caller-zeroed) += u[e] * (use_tanh ? tanh(sc[e]) * coords_range : sc[e]) / div
 *   bwd: gu [E][3], gsc [E] from gstep [N][3] */
int geoldm_train_coord2diff_fwd(int n_edge, const float* x, const int* edge_i, const int* edge_j, float norm_constant,
                                float* r, float* u, void* stream);
int geoldm_train_coord2diff_bwd(int n_edge, const float* x, const int* edge_i, const int* edge_j, float norm_constant,
                                const float* gr, const float* gu, float* gx, void* stream);
int geoldm_train_coord_step_fwd(int n_edge, const float* u, const float* sc, const int* edge_i, int use_tanh,
                                float coords_range, float div, float* step, void* stream);
int geoldm_train_coord_step_bwd(int n_edge, const float* u, const float* sc, const int* edge_i, int use_tanh,
                                float coords_range, float div, const float* gstep, float* gu, float* gsc, void* stream);
/* number of thread blocks the backward edge kernels launch for n_edge edges (sizes dbw_scratch) */
int geoldm_train_bwd_blocks(int n_edge);
/* ---- evaluation side: bond-order stability of a ragged batch of molecules.
 * Replaces the Python double loop of qm9/analyze.py:209-245 (check_stability) + qm9/bond_analyze.py:101-146.
 * x [N][3] fp32, atom_type [N] in [0, n_types), mol_off [n_mol+1]; thr [3][n_types][n_types] = single/double/triple
 * thresholds in pm incl. margins for the ORDERED pair (first, second), negative = no entry; allowed[t] = bitmask of
 * allowed valences; sorted_pair 0: (type_i, type_j) with i < j (QM9), 1: pair sorted by type index (GEOM).
 * Outputs: nr_bonds [N] (may be NULL), n_stable [n_mol] = number of atoms with an allowed valence. */
int geoldm_stability(int n_mol, const int* mol_off, const float* x, const int* atom_type, int n_types, const float* thr,
                     const int* allowed, int sorted_pair, int* nr_bonds, int* n_stable, void* stream);
/* debug (GEOLDM_TC_DEBUG & 32): read+reset cycle counters of the MMA-issuing thread of CTA 0:
 * {total, wait acc_empty, wait a_full, wait w_full, launches, tiles, 0, 0}; synchronises the device */
int geoldm_tc_read_stats(unsigned long long* host_out);
/* same for the fp16-split kernels (16 counters: MMA thread, one producer thread, one epilogue thread of CTA 0) */
int geoldm_tc16_read_stats(unsigned long long* host_out);
/* Philox4x32-10 standard normals (Box-Muller), the sampler's noise stream exposed for tests:
 * out[4*d+e] = e-th normal of block counter=(d, node, blk, seed>>32), key=(mol_id, (uint32)seed), i.e. what
 * geoldm_sampler_update draws for draw index d, node `node` of molecule `mol_id`, columns 4*blk+e. */
int geoldm_philox_normal(uint64_t seed, uint64_t mol_id, uint32_t node, uint32_t blk, float* out, int n,
                         void* stream);
/* EnHierarchicalVAE.decode tail (en_diffusion.py:1028-1033): one_hot = onehot(argmax(h[:, :n_cat]), n_classes),
 * charges = round(h[:, charge_col]) (charge_col < 0: none).  Outputs int64, ragged [N][n_classes] / [N]. */
int geoldm_decode(int n_node, const float* h, int h_dim, int n_cat, int n_classes, int charge_col,
                  long long* one_hot, long long* charges, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOLDM_B200_H */

// C-ABI entry points: error plumbing, EGNN.forward orchestration, unit-level wrappers.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace geoldm {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  float *h, *h2, *t1, *agg, *pq, *xa, *xb, *xagg, *dx;
  float *r_edge, *d0_edge;   // [E] squared distances of the current / entry coordinates (tensor-core edge kernels)
  float *u_edge;             // [E][4] normalised coordinate differences of the current coordinates
  size_t bytes;
};
static Workspace carve(void* base, int n_node, int n_edge, int H) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(n_floats * sizeof(float), 256);
    return p;
  };
  const size_t nh = (size_t)n_node * H;
  w.h = take(nh); w.h2 = take(nh); w.t1 = take(nh); w.agg = take(nh); w.pq = take(4 * nh);
  w.xa = take((size_t)3 * n_node); w.xb = take((size_t)3 * n_node); w.xagg = take((size_t)3 * n_node); w.dx = take((size_t)3 * n_node);
  w.r_edge = take((size_t)n_edge); w.d0_edge = take((size_t)n_edge); w.u_edge = take((size_t)4 * n_edge);
  w.bytes = off;
  return w;
}

static int edge_dispatch(const geoldm_egnn_config& cfg, const geoldm_edge_mlp& w, const geoldm_batch& b, bool equiv,
                         const float* pq, int pq_ld, const float* x, const float* x0, float* out, cudaStream_t st,
                         const float* r_edge = nullptr, const float* d0_edge = nullptr, const float* u_edge = nullptr) {
  if (cfg.mma_mode == GEOLDM_MMA_FP32_SIMT) {
    GEOLDM_REQUIRE(pq_ld == 2 * cfg.hidden_nf, "edge_simt expects a [N][2H] projection buffer");
    return launch_edge_simt(cfg, w, b, equiv, pq, x, x0, out, st);
  }
  if (cfg.mma_mode == GEOLDM_MMA_3XF16)
    return launch_edge_tc16(cfg, w, b, equiv, pq, pq_ld, x, x0, r_edge, d0_edge, u_edge, out, st);
  return launch_edge_tc(cfg, w, b, equiv, pq, pq_ld, x, x0, out, st);
}
}  // namespace geoldm

using namespace geoldm;

extern "C" {

int geoldm_abi_version(void) { return GEOLDM_ABI_VERSION; }
int geoldm_node_chain(int H, const float* h, const float* agg, float agg_div, const void* pack1, const float* b1,
                      const void* pack2, const float* b2, const void* pack3, const float* b3, int n_blocks3, float* h_out,
                      float* pq_out, float* agg_zero, int m, void* stream) {
  return geoldm::launch_node_chain16(H, h, agg, agg_div, pack1, b1, pack2, b2, pack3, b3, n_blocks3, h_out, pq_out, agg_zero, m,
                                     (cudaStream_t)stream);
}
const char* geoldm_last_error(void) { return g_err; }

size_t geoldm_egnn_workspace_bytes(const geoldm_egnn_config* cfg, int n_node, int n_edge) {
  return carve(nullptr, n_node, n_edge, cfg->hidden_nf).bytes;
}

static int check_cfg(const geoldm_egnn_config* cfg) {
  GEOLDM_REQUIRE(cfg->n_layers >= 1 && cfg->n_layers <= GEOLDM_MAX_LAYERS, "n_layers %d not in [1,%d]", cfg->n_layers,
                 GEOLDM_MAX_LAYERS);
  GEOLDM_REQUIRE(cfg->inv_sublayers >= 1 && cfg->inv_sublayers <= GEOLDM_MAX_SUBLAYERS, "inv_sublayers %d not in [1,%d]",
                 cfg->inv_sublayers, GEOLDM_MAX_SUBLAYERS);
  GEOLDM_REQUIRE(cfg->hidden_nf % 32 == 0, "hidden_nf %d must be a multiple of 32", cfg->hidden_nf);
  GEOLDM_REQUIRE(cfg->agg_div != 0.f, "agg_div must be non-zero");
  GEOLDM_REQUIRE(cfg->mma_mode >= 0 && cfg->mma_mode <= 4, "bad mma_mode %d", cfg->mma_mode);
  return 0;
}

int geoldm_egnn_forward(const geoldm_egnn_config* cfg, const geoldm_egnn_weights* w, const geoldm_batch* b,
                        const float* h_in, const float* x_in, float* h_out, float* x_out, float* dx_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = b->n_node, H = cfg->hidden_nf;
  const bool tcore = cfg->mma_mode != GEOLDM_MMA_FP32_SIMT;
  const int terms = cfg->mma_mode == GEOLDM_MMA_TF32 ? 1 : cfg->mma_mode == GEOLDM_MMA_3XF16 ? 16 : 3;
  GEOLDM_REQUIRE(cfg->mma_mode != GEOLDM_MMA_BF16, "mma_mode bf16 is not implemented yet");
  if (N == 0) return 0;
  Workspace ws = carve(workspace, N, b->n_edge, H);
  GEOLDM_REQUIRE(workspace && workspace_bytes >= ws.bytes, "egnn_forward: workspace %zu < %zu bytes", workspace_bytes,
                 ws.bytes);
  int rc;
  if ((rc = launch_embed(N, H, cfg->in_node_nf, h_in, w->emb_w, w->emb_b, ws.h, st))) return rc;
  const float* x_cur = x_in;  // EGNN-entry coordinates stay in x_in (d0 is recomputed from them)
  float* x_bufs[2] = {ws.xa, ws.xb};
  int xi = 0;
  float *h = ws.h, *h2 = ws.h2;
  // P|Q of the upcoming gcl_0: [N][pq_ld] at pq_next; produced either stand-alone or by the previous block's fused
  // 4-column-block projection (equiv P|Q in columns [0,2H), next gcl_0 P|Q in [2H,4H))
  const float* pq_next = nullptr;
  int pq_next_ld = 2 * H;
  // fp16-split kernels read the per-edge squared distances instead of recomputing them in every producer thread
  const bool pre_dist = cfg->mma_mode == GEOLDM_MMA_3XF16;
  const float* r_edge = pre_dist ? ws.r_edge : nullptr;
  const float* d0_edge = pre_dist ? ws.d0_edge : nullptr;
  const float* u_edge = pre_dist ? ws.u_edge : nullptr;
  if (pre_dist && (rc = launch_edge_dist(*b, x_in, ws.d0_edge, ws.u_edge, cfg->norm_constant, st))) return rc;
  // segment-sum targets: the fp16-split path clears them ONCE here; afterwards their consumers (the residual epilogue
  // of node_mlp.2 for agg, the coordinate update for xagg) hand them back zeroed - 2 memset nodes per forward instead
  // of 2 per block.  The other modes clear them before every edge kernel.
  const bool self_clean = cfg->mma_mode == GEOLDM_MMA_3XF16;
  if (self_clean) {
    cudaError_t e1 = cudaMemsetAsync(ws.agg, 0, (size_t)N * H * sizeof(float), st);
    cudaError_t e2 = cudaMemsetAsync(ws.xagg, 0, (size_t)3 * N * sizeof(float), st);
    GEOLDM_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, "egnn_forward: memset failed: %s",
                   cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  }
  // fp16-split path, opt-in (GEOLDM_TC_CHAIN=1): node_mlp.0 -> SiLU -> node_mlp.2 -> + h -> the projections that read the
  // new h as ONE persistent launch per GCL (chain16_kernel: intermediates stay in shared memory as operand images).
  // Parity-green, but MEASURED SLOWER than the three dense launches at the bench size (105.7 vs 99.9 us per block at
  // 22 576 rows: 89 tile pairs on 74 CTA pairs make two full rounds of the whole chain, and a tile's phases run
  // back to back with their epilogue latencies exposed; profiles/README.md), so the dense launches stay the default.
  static int chain_env = -1;
  if (chain_env < 0) { const char* e = getenv("GEOLDM_TC_CHAIN"); chain_env = e ? atoi(e) : 0; }
  const bool chain = chain_env && cfg->mma_mode == GEOLDM_MMA_3XF16;
  for (int l = 0; l < cfg->n_layers; ++l) {
    const geoldm_block& blk = w->block[l];
    bool pq_ready = false;          // the chain kernel has already produced the projections the next consumer needs
    bool equiv_pq_ready = false;
    int equiv_pq_ld = 2 * H;
    if (pre_dist) {
      if (l == 0) r_edge = ws.d0_edge;                    // x == x_in in the first block
      else { if ((rc = launch_edge_dist(*b, x_cur, ws.r_edge, ws.u_edge, cfg->norm_constant, st))) return rc; r_edge = ws.r_edge; }
    }
    for (int s = 0; s < cfg->inv_sublayers; ++s) {
      const geoldm_gcl& g = blk.gcl[s];
      const float* pq = ws.pq;
      int pq_ld = 2 * H;
      if (s == 0 && pq_next) {
        pq = pq_next; pq_ld = pq_next_ld;
      } else if (pq_ready) {
        pq_ready = false;             // written into ws.pq by the previous sublayer's chain kernel
      } else if (tcore) {
        if ((rc = launch_linear_tc(H, terms, h, H, nullptr, 0, 1.f, g.edge.tc_pack_pq, 2, g.edge.pq_b, nullptr, 0, ws.pq, N, st))) return rc;
      } else {
        if ((rc = launch_linear(h, H, nullptr, 0, 1.f, g.edge.pq_wt, g.edge.pq_b, nullptr, 0, ws.pq, N, 2 * H, st))) return rc;
      }
      if (!self_clean) {
        cudaError_t e = cudaMemsetAsync(ws.agg, 0, (size_t)N * H * sizeof(float), st);
        GEOLDM_REQUIRE(e == cudaSuccess, "egnn_forward: memset failed: %s", cudaGetErrorString(e));
      }
      if ((rc = edge_dispatch(*cfg, g.edge, *b, false, pq, pq_ld, x_cur, x_in, ws.agg, st, r_edge, d0_edge))) return rc;
      if (chain) {
        const void* pack3; const float* b3; int n_pb;
        if (s + 1 < cfg->inv_sublayers) {                       // next sublayer's GCL projections
          pack3 = blk.gcl[s + 1].edge.tc_pack_pq; b3 = blk.gcl[s + 1].edge.pq_b; n_pb = 2; pq_ready = true;
        } else if (blk.tc_pack_pq4 && l + 1 < cfg->n_layers) {  // this block's coord_mlp + next block's gcl_0
          pack3 = blk.tc_pack_pq4; b3 = blk.pq4_b; n_pb = 4; equiv_pq_ready = true; equiv_pq_ld = 4 * H;
        } else {                                                // this block's coord_mlp only
          pack3 = blk.equiv.tc_pack_pq; b3 = blk.equiv.pq_b; n_pb = 2; equiv_pq_ready = true; equiv_pq_ld = 2 * H;
        }
        if ((rc = launch_node_chain16(H, h, ws.agg, cfg->agg_div, g.tc_pack_node1, g.node_b1, g.tc_pack_node2, g.node_b2,
                                      pack3, b3, n_pb, h2, ws.pq, ws.agg, N, st))) return rc;
      } else if (tcore) {
        if ((rc = launch_linear_tc(H, terms, h, H, ws.agg, H, cfg->agg_div, g.tc_pack_node1, 1, g.node_b1, nullptr, 1, ws.t1, N, st))) return rc;
        if ((rc = launch_linear_tc(H, terms, ws.t1, H, nullptr, 0, 1.f, g.tc_pack_node2, 1, g.node_b2, h, 2, h2, N, st,
                                   self_clean ? ws.agg : nullptr))) return rc;
      } else {
        if ((rc = launch_linear(h, H, ws.agg, H, cfg->agg_div, g.node_w1t, g.node_b1, nullptr, 1, ws.t1, N, H, st))) return rc;
        if ((rc = launch_linear(ws.t1, H, nullptr, 0, 1.f, g.node_w2t, g.node_b2, h, 2, h2, N, H, st))) return rc;
      }
      float* tmp = h; h = h2; h2 = tmp;
    }
    const geoldm_edge_mlp& e = blk.equiv;
    int pq_ld = 2 * H;
    pq_next = nullptr;
    if (equiv_pq_ready) {
      pq_ld = equiv_pq_ld;
      if (pq_ld == 4 * H) { pq_next = ws.pq + 2 * H; pq_next_ld = 4 * H; }
    } else if (tcore && blk.tc_pack_pq4 && l + 1 < cfg->n_layers) {
      if ((rc = launch_linear_tc(H, terms, h, H, nullptr, 0, 1.f, blk.tc_pack_pq4, 4, blk.pq4_b, nullptr, 0, ws.pq, N, st))) return rc;
      pq_ld = 4 * H;
      pq_next = ws.pq + 2 * H;
      pq_next_ld = 4 * H;
    } else if (tcore) {
      if ((rc = launch_linear_tc(H, terms, h, H, nullptr, 0, 1.f, e.tc_pack_pq, 2, e.pq_b, nullptr, 0, ws.pq, N, st))) return rc;
    } else {
      if ((rc = launch_linear(h, H, nullptr, 0, 1.f, e.pq_wt, e.pq_b, nullptr, 0, ws.pq, N, 2 * H, st))) return rc;
    }
    if (!self_clean) {
      cudaError_t em = cudaMemsetAsync(ws.xagg, 0, (size_t)3 * N * sizeof(float), st);
      GEOLDM_REQUIRE(em == cudaSuccess, "egnn_forward: memset failed: %s", cudaGetErrorString(em));
    }
    if ((rc = edge_dispatch(*cfg, e, *b, true, ws.pq, pq_ld, x_cur, x_in, ws.xagg, st, r_edge, d0_edge, u_edge))) return rc;
    const bool last = (l + 1 == cfg->n_layers);
    float* x_next = last ? x_out : x_bufs[xi];
    float* dx_next = (last && dx_out) ? dx_out : ws.dx;   // dx is updated in place (elementwise)
    if ((rc = launch_coord_update(3 * N, x_in, l == 0 ? nullptr : ws.dx, ws.xagg, cfg->agg_div, dx_next, x_next, st,
                                  self_clean)))
      return rc;
    x_cur = x_next;
    xi ^= 1;
  }
  if ((rc = launch_outproj(N, H, cfg->out_node_nf, h, w->out_w, w->out_b, h_out, st))) return rc;
  GEOLDM_CHECK_LAUNCH("egnn_forward");
  return 0;
}

int geoldm_edge_gcl(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b, const float* pq,
                    const float* x, const float* x0, float* agg, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  return edge_dispatch(*cfg, *w, *b, false, pq, 2 * cfg->hidden_nf, x, x0, agg, (cudaStream_t)stream);
}
int geoldm_edge_equiv(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b, const float* pq,
                      const float* x, const float* x0, float* xagg, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  return edge_dispatch(*cfg, *w, *b, true, pq, 2 * cfg->hidden_nf, x, x0, xagg, (cudaStream_t)stream);
}
int geoldm_edge_dist(const geoldm_batch* b, const float* x, float* r_out, float* u_out, float norm_constant,
                     void* stream) {
  return launch_edge_dist(*b, x, r_out, u_out, norm_constant, (cudaStream_t)stream);
}
int geoldm_edge_gcl_pre(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b, const float* pq,
                        int pq_ld, const float* r_edge, const float* d0_edge, float* agg, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  GEOLDM_REQUIRE(cfg->mma_mode == GEOLDM_MMA_3XF16 && r_edge && d0_edge, "edge_gcl_pre: mma_mode 3xf16 with r_edge/d0_edge");
  return edge_dispatch(*cfg, *w, *b, false, pq, pq_ld, nullptr, nullptr, agg, (cudaStream_t)stream, r_edge, d0_edge);
}
int geoldm_edge_equiv_pre(const geoldm_egnn_config* cfg, const geoldm_edge_mlp* w, const geoldm_batch* b,
                          const float* pq, int pq_ld, const float* r_edge, const float* d0_edge, const float* u_edge,
                          float* xagg, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  GEOLDM_REQUIRE(cfg->mma_mode == GEOLDM_MMA_3XF16 && r_edge && d0_edge && u_edge,
                 "edge_equiv_pre: mma_mode 3xf16 with r_edge/d0_edge/u_edge");
  return edge_dispatch(*cfg, *w, *b, true, pq, pq_ld, nullptr, nullptr, xagg, (cudaStream_t)stream, r_edge, d0_edge,
                       u_edge);
}
int geoldm_linear(const float* a1, int k1, const float* a2, int k2, float a2_div, const float* wt, const float* bias,
                  const float* res, int epi, float* out, int m, int n, int mma_mode, void* stream) {
  (void)mma_mode;
  return launch_linear(a1, k1, a2, k2, a2_div, wt, bias, res, epi, out, m, n, (cudaStream_t)stream);
}


int geoldm_linear_tc(int H, int terms, const float* a1, int k1, const float* a2, int k2, float a2_div,
                     const void* w_pack, int n_blocks, const float* bias, const float* res, int epi, float* out, int m,
                     void* stream) {
  return launch_linear_tc(H, terms, a1, k1, a2, k2, a2_div, w_pack, n_blocks, bias, res, epi, out, m,
                          (cudaStream_t)stream);
}
int geoldm_tc_selftest(int H, int terms, const float* a, const int* src_row, const int* tile_row, int n_tile, int n_rows,
                       const void* w_pack, float* out, void* stream) {
  return launch_tc_selftest(H, terms, a, src_row, tile_row, n_tile, n_rows, w_pack, out, (cudaStream_t)stream);
}

}  // extern "C"

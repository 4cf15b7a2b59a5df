// Small HBM-bound kernels around the EGNN: feature assembly, embeddings, coordinate update,
// velocity / centre-of-mass projection, the ancestral-sampling update and the Philox noise source.
// Reference: egnn/models.py:56-76,80-113; egnn/egnn_new.py:172-173,189,194-196;
//            equivariant_diffusion/utils.py:31-38,107-116,137-140;
//            equivariant_diffusion/en_diffusion.py:716-760,1099-1122.
#include "common.cuh"

namespace geoldm {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"), Box-Muller normals.
// ---------------------------------------------------------------------------------------------
struct Philox {
  uint32_t c[4];
};
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// four standard normals from one Philox block
__device__ __forceinline__ void philox_normal4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, float (&out)[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
  philox4x32_10(c, k0, k1);
  const float s = 5.9604644775390625e-8f;  // 2^-24
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float u1 = ((float)(c[2 * p] >> 8) + 0.5f) * s;      // (0,1)
    const float u2 = ((float)(c[2 * p + 1] >> 8) + 0.5f) * s;
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    out[2 * p] = r * cs;
    out[2 * p + 1] = r * sn;
  }
}

__global__ void philox_normal_kernel(uint32_t k0, uint32_t k1, uint32_t c1, uint32_t c2, uint32_t c3, float* out,
                                     int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * b >= n) return;
  float v[4];
  philox_normal4((uint32_t)b, c1, c2, c3, k0, k1, v);
  for (int e = 0; e < 4 && 4 * b + e < n; ++e) out[4 * b + e] = v[e];
}

// ---------------------------------------------------------------------------------------------
// feature assembly + embeddings
// ---------------------------------------------------------------------------------------------
__global__ void prep_kernel(int n_node, const int* __restrict__ node_mol, const int* __restrict__ node_src,
                            const float* __restrict__ xh, int xh_dim, const float* __restrict__ t_mol,
                            const float* __restrict__ t_table, const int* __restrict__ step_idx,
                            const float* __restrict__ context, int ctx_nf, int condition_time,
                            float* __restrict__ h_in, int in_nf, float* __restrict__ x) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_node) return;
  const int src = node_src ? node_src[k] : k;
  const float* row = xh + (size_t)src * xh_dim;
  x[3 * k] = row[0]; x[3 * k + 1] = row[1]; x[3 * k + 2] = row[2];
  const int lat = xh_dim - 3;
  float* h = h_in + (size_t)k * in_nf;
  for (int f = 0; f < lat; ++f) h[f] = row[3 + f];
  int o = lat;
  if (condition_time) {
    h[o++] = t_mol ? t_mol[node_mol[k]] : t_table[4 * (*step_idx) + 3];
  }
  for (int c = 0; c < ctx_nf; ++c) h[o + c] = context[(size_t)src * ctx_nf + c];
}

// h[N][H] = h_in[N][F] * W^T + b,  W [H][F] (PyTorch layout), F small
__global__ void embed_kernel(int n_node, int H, int F, const float* __restrict__ h_in, const float* __restrict__ w,
                             const float* __restrict__ b, float* __restrict__ h) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_node * H) return;
  const int node = (int)(idx / H), f = (int)(idx % H);
  float acc = 0.f;
  for (int k = 0; k < F; ++k) acc = fmaf(h_in[(size_t)node * F + k], w[(size_t)f * F + k], acc);
  h[idx] = acc + b[f];
}

// out[N][Fo] = h[N][H] * W^T + b, W [Fo][H]; one warp per (node, chunk of OUTPROJ_FC outputs), the node's row kept in
// registers across the outputs.  (The VAE encoder's EGNN has Fo = hidden_nf outputs: one warp walking all of them in turn
// took 0.25 ms per launch at 1158 nodes.)
constexpr int OUTPROJ_FC = 16;
__global__ void outproj_kernel(int n_node, int H, int Fo, const float* __restrict__ h, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ out) {
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= n_node) return;
  const float* hr = h + (size_t)node * H;
  float hv[8];                                   // the first 256 columns
#pragma unroll
  for (int j = 0; j < 8; ++j) hv[j] = (lane + 32 * j < H) ? hr[lane + 32 * j] : 0.f;
  const int f0 = blockIdx.y * OUTPROJ_FC, f1 = min(Fo, f0 + OUTPROJ_FC);
  for (int f = f0; f < f1; ++f) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (lane + 32 * j < H) acc = fmaf(hv[j], w[(size_t)f * H + lane + 32 * j], acc);
    for (int k = lane + 256; k < H; k += 32) acc = fmaf(hr[k], w[(size_t)f * H + k], acc);   // H > 256: rest of the row from memory
    acc = warp_sum(acc);
    if (lane == 0) out[(size_t)node * Fo + f] = acc + b[f];
  }
}

// EquivariantUpdate: coord + agg (egnn_new.py:95-98), kept as x0 + accumulated displacement:
//   dx_next = dx + xagg / div ;  x_next = x0 + dx_next
// The reference adds into x block after block and finally subtracts x0 again (egnn/models.py:80), which costs
// ~ulp(|x|) of cancellation noise on a velocity that is 30x smaller than x; carrying the displacement keeps
// the velocity at full precision (closer to the exact result, hence also closer to the reference: errors of
// two fp32 evaluations add in quadrature).
__global__ void coord_update_kernel(int n, const float* __restrict__ x0, const float* __restrict__ dx,
                                    float* __restrict__ xagg, float div, float* __restrict__ dx_next,
                                    float* __restrict__ x_next, bool zero_xagg) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float a = xagg[k];
  if (zero_xagg) xagg[k] = 0.f;          // consumed: ready for the next block's segment sums (no memset node)
  if (div != 1.0f) a = __fdiv_rn(a, div);
  const float d = dx ? __fadd_rn(dx[k], a) : a;
  dx_next[k] = d;
  x_next[k] = __fadd_rn(x0[k], d);
}

// ---------------------------------------------------------------------------------------------
// velocity, NaN guard, CoM projection
// ---------------------------------------------------------------------------------------------
__global__ void finish_a_kernel(int n3, const float* __restrict__ vel, int* __restrict__ nan_flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n3) return;
  if (isnan(vel[k])) atomicOr(nan_flag, 1);
}

constexpr int MAX_PER_LANE = 8;  // molecules up to 256 atoms

// one warp per molecule
__global__ void finish_b_kernel(int n_mol, const int* __restrict__ mol_off, const int* __restrict__ node_src,
                                const float* __restrict__ vel, const float* __restrict__ h_out, int h_stride,
                                int h_keep, const int* __restrict__ nan_flag, float* __restrict__ out,
                                int out_dim) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= n_mol) return;
  const int n0 = mol_off[m], n = mol_off[m + 1] - n0;
  const bool bad = (*nan_flag) != 0;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int a = lane; a < n; a += 32) {
    const float* v = vel + 3 * (size_t)(n0 + a);
    if (!bad) { sx += v[0]; sy += v[1]; sz += v[2]; }
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  const float fn = (float)n;
  const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
  for (int a = lane; a < n; a += 32) {
    const int k = n0 + a;
    const int dst = node_src ? node_src[k] : k;
    float* o = out + (size_t)dst * out_dim;
    const float* v = vel + 3 * (size_t)k;
    o[0] = bad ? 0.f : __fsub_rn(v[0], mx);
    o[1] = bad ? 0.f : __fsub_rn(v[1], my);
    o[2] = bad ? 0.f : __fsub_rn(v[2], mz);
    for (int f = 0; f < h_keep; ++f) o[3 + f] = h_out[(size_t)k * h_stride + f];
  }
}

// ---------------------------------------------------------------------------------------------
// ancestral sampling update  (one warp per molecule, dim = 3 + latent <= 8)
// ---------------------------------------------------------------------------------------------
__global__ void sampler_update_kernel(int n_mol, const int* __restrict__ mol_off, int mode,
                                      const float* __restrict__ coef, const int* __restrict__ step_idx,
                                      const float* __restrict__ z, const float* __restrict__ eps_hat,
                                      const float* __restrict__ noise, size_t noise_stride, int dim,
                                      uint32_t seed_lo, uint32_t seed_hi,
                                      const int64_t* __restrict__ mol_id, const int* __restrict__ draw_idx,
                                      float* __restrict__ z_out) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= n_mol) return;
  const int n0 = mol_off[m], n = mol_off[m + 1] - n0;
  const int step = step_idx ? *step_idx : 0;
  const float c0 = coef ? coef[4 * step] : 1.f, c1 = coef ? coef[4 * step + 1] : 0.f,
              c2 = coef ? coef[4 * step + 2] : 1.f;
  const uint32_t gid = mol_id ? (uint32_t)mol_id[m] : (uint32_t)m;
  const uint32_t draw = draw_idx ? (uint32_t)(*draw_idx) : 0u;
  if (noise) noise += (size_t)draw * noise_stride;

  float nz[MAX_PER_LANE][8];
#pragma unroll
  for (int q = 0; q < MAX_PER_LANE; ++q)
#pragma unroll
    for (int d = 0; d < 8; ++d) nz[q][d] = 0.f;
  // 1) noise (injected or Philox), masked to real nodes by construction
  float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
  for (int q = 0; q < MAX_PER_LANE; ++q) {
    const int a = lane + 32 * q;
    if (a < n) {
      if (noise) {
        for (int d = 0; d < dim; ++d) nz[q][d] = noise[(size_t)(n0 + a) * dim + d];
      } else {
        float v[4];
        philox_normal4(draw, (uint32_t)a, 0u, seed_hi, gid, seed_lo, v);
        nz[q][0] = v[0]; nz[q][1] = v[1]; nz[q][2] = v[2]; nz[q][3] = v[3];
        if (dim > 4) {
          philox_normal4(draw, (uint32_t)a, 1u, seed_hi, gid, seed_lo, v);
          nz[q][4] = v[0]; nz[q][5] = v[1]; nz[q][6] = v[2]; nz[q][7] = v[3];
        }
      }
      sx += nz[q][0]; sy += nz[q][1]; sz += nz[q][2];
    }
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  const float fn = (float)n;
  {
    const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
#pragma unroll
    for (int q = 0; q < MAX_PER_LANE; ++q) {
      nz[q][0] = __fsub_rn(nz[q][0], mx); nz[q][1] = __fsub_rn(nz[q][1], my); nz[q][2] = __fsub_rn(nz[q][2], mz);
    }
  }
  // 2) update
  sx = sy = sz = 0.f;
#pragma unroll
  for (int q = 0; q < MAX_PER_LANE; ++q) {
    const int a = lane + 32 * q;
    if (a < n) {
      for (int d = 0; d < dim; ++d) {
        const size_t idx = (size_t)(n0 + a) * dim + d;
        float o;
        if (mode == 0) {         // mu = z/alpha_ts - c_eps*eps ; z_s = mu + sigma*noise   (en_diffusion.py:733-739)
          const float mu = __fsub_rn(__fdiv_rn(z[idx], c0), __fmul_rn(c1, eps_hat[idx]));
          o = __fadd_rn(mu, __fmul_rn(c2, nz[q][d]));
        } else if (mode == 1) {  // mu = (1/alpha0) * (z0 - sigma0*eps) ; xh = mu + sigma_x*noise   (:446, :1108-1109)
          const float mu = __fmul_rn(c0, __fsub_rn(z[idx], __fmul_rn(c1, eps_hat[idx])));
          o = __fadd_rn(mu, __fmul_rn(c2, nz[q][d]));
        } else {
          o = nz[q][d];
        }
        nz[q][d] = o;
      }
      sx += nz[q][0]; sy += nz[q][1]; sz += nz[q][2];
    }
  }
  // 3) CoM projection of the x part (mode 0 only, :742-746); modes 1/2 are written as is
  float mx = 0.f, my = 0.f, mz = 0.f;
  if (mode == 0) {
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    mx = __fdiv_rn(sx, fn); my = __fdiv_rn(sy, fn); mz = __fdiv_rn(sz, fn);
  }
#pragma unroll
  for (int q = 0; q < MAX_PER_LANE; ++q) {
    const int a = lane + 32 * q;
    if (a < n) {
      float* o = z_out + (size_t)(n0 + a) * dim;
      o[0] = __fsub_rn(nz[q][0], mx); o[1] = __fsub_rn(nz[q][1], my); o[2] = __fsub_rn(nz[q][2], mz);
      for (int d = 3; d < dim; ++d) o[d] = nz[q][d];
    }
  }
}

__global__ void advance_kernel(int* step_idx, int d_step, int* draw_idx, int d_draw) {
  if (step_idx) *step_idx += d_step;
  if (draw_idx) *draw_idx += d_draw;
}

// decode: argmax one-hot over h_cat = xh[:, 3:-1] quirk handled by the caller via (cat_off, n_cat)
__global__ void decode_kernel(int n_node, const float* __restrict__ h, int h_dim, int n_cat, int n_classes,
                              int charge_col, long long* __restrict__ one_hot, long long* __restrict__ charges) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_node) return;
  const float* r = h + (size_t)k * h_dim;
  int best = 0;
  float bv = r[0];
  for (int c = 1; c < n_cat; ++c) {
    // torch.argmax returns the first maximal index; NaN is treated as maximal
    if (r[c] > bv || (isnan(r[c]) && !isnan(bv))) { bv = r[c]; best = c; }
  }
  for (int c = 0; c < n_classes; ++c) one_hot[(size_t)k * n_classes + c] = (c == best) ? 1 : 0;
  if (charge_col >= 0) charges[k] = (long long)rintf(r[charge_col]);
}

}  // namespace geoldm

using namespace geoldm;

extern "C" {

int geoldm_dynamics_prep(const geoldm_batch* b, const int* node_src, const float* xh, int xh_dim, const float* t_mol,
                         const float* t_table, const int* step_idx_dev, const float* context, int ctx_nf,
                         int condition_time, float* h_in, int in_node_nf, float* x, void* stream) {
  GEOLDM_REQUIRE(xh_dim >= 3, "prep: xh_dim %d < 3", xh_dim);
  GEOLDM_REQUIRE(in_node_nf == (xh_dim - 3) + (condition_time ? 1 : 0) + ctx_nf,
                 "prep: in_node_nf %d != latent %d + time %d + ctx %d", in_node_nf, xh_dim - 3, condition_time, ctx_nf);
  GEOLDM_REQUIRE(!condition_time || t_mol || (t_table && step_idx_dev), "prep: no time source");
  GEOLDM_REQUIRE(ctx_nf == 0 || context, "prep: ctx_nf %d but context is NULL", ctx_nf);
  if (b->n_node == 0) return 0;
  prep_kernel<<<(b->n_node + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      b->n_node, b->node_mol, node_src, xh, xh_dim, t_mol, t_table, step_idx_dev, context, ctx_nf, condition_time,
      h_in, in_node_nf, x);
  GEOLDM_CHECK_LAUNCH("prep_kernel");
  return 0;
}

int geoldm_dynamics_finish_a(const geoldm_batch* b, const float* vel, int* nan_flag, void* stream) {
  const int n3 = 3 * b->n_node;
  if (n3 == 0) return 0;
  finish_a_kernel<<<(n3 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n3, vel, nan_flag);
  GEOLDM_CHECK_LAUNCH("finish_a_kernel");
  return 0;
}

int geoldm_dynamics_finish_b(const geoldm_batch* b, const int* node_src, const float* vel, const float* h_out,
                             int h_stride, int h_keep, const int* nan_flag, float* out, int out_dim, void* stream) {
  GEOLDM_REQUIRE(out_dim >= 3 + h_keep && h_keep <= h_stride, "finish_b: out_dim %d, h_keep %d, h_stride %d", out_dim,
                 h_keep, h_stride);
  if (b->n_mol == 0) return 0;
  finish_b_kernel<<<(b->n_mol * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      b->n_mol, b->mol_off, node_src, vel, h_out, h_stride, h_keep, nan_flag, out, out_dim);
  GEOLDM_CHECK_LAUNCH("finish_b_kernel");
  return 0;
}

int geoldm_sampler_update(const geoldm_batch* b, int mode, const float* coef, const int* step_idx_dev, const float* z,
                          const float* eps_hat, const float* noise, size_t noise_stride, int dim, uint64_t seed,
                          const int64_t* mol_id,
                          const int* draw_idx_dev, float* z_out, void* stream) {
  GEOLDM_REQUIRE(dim >= 3 && dim <= 8, "sampler_update: dim %d not in [3,8]", dim);
  GEOLDM_REQUIRE(mode >= 0 && mode <= 2, "sampler_update: mode %d", mode);
  GEOLDM_REQUIRE(mode == 2 || (z && eps_hat && coef), "sampler_update: mode %d needs z, eps_hat, coef", mode);
  if (b->n_mol == 0) return 0;
  sampler_update_kernel<<<(b->n_mol * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      b->n_mol, b->mol_off, mode, coef, step_idx_dev, z, eps_hat, noise, noise_stride, dim, (uint32_t)seed,
      (uint32_t)(seed >> 32),
      mol_id, draw_idx_dev, z_out);
  GEOLDM_CHECK_LAUNCH("sampler_update_kernel");
  return 0;
}

int geoldm_sampler_advance(int* step_idx_dev, int d_step, int* draw_idx_dev, int d_draw, void* stream) {
  advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_idx_dev, d_step, draw_idx_dev, d_draw);
  GEOLDM_CHECK_LAUNCH("advance_kernel");
  return 0;
}

int geoldm_philox_normal(uint64_t seed, uint64_t key_hi_lo, uint32_t c1, uint32_t c2, float* out, int n,
                         void* stream) {
  if (n == 0) return 0;
  const int nb = (n + 3) / 4;
  philox_normal_kernel<<<(nb + 127) / 128, 128, 0, (cudaStream_t)stream>>>((uint32_t)key_hi_lo, (uint32_t)seed, c1, c2,
                                                                            (uint32_t)(seed >> 32), out, n);
  GEOLDM_CHECK_LAUNCH("philox_normal_kernel");
  return 0;
}

int geoldm_decode(int n_node, const float* h, int h_dim, int n_cat, int n_classes, int charge_col, long long* one_hot,
                  long long* charges, void* stream) {
  GEOLDM_REQUIRE(n_cat >= 1 && n_cat <= h_dim && n_cat <= n_classes, "decode: n_cat %d h_dim %d n_classes %d", n_cat,
                 h_dim, n_classes);
  if (n_node == 0) return 0;
  decode_kernel<<<(n_node + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_node, h, h_dim, n_cat, n_classes, charge_col,
                                                                        one_hot, charges);
  GEOLDM_CHECK_LAUNCH("decode_kernel");
  return 0;
}

}  // extern "C"

namespace geoldm {
int launch_embed(int n_node, int H, int F, const float* h_in, const float* w, const float* b, float* h,
                 cudaStream_t st) {
  const size_t tot = (size_t)n_node * H;
  if (tot == 0) return 0;
  embed_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n_node, H, F, h_in, w, b, h);
  GEOLDM_CHECK_LAUNCH("embed_kernel");
  return 0;
}
int launch_outproj(int n_node, int H, int Fo, const float* h, const float* w, const float* b, float* out,
                   cudaStream_t st) {
  if (n_node == 0) return 0;
  outproj_kernel<<<dim3((n_node * 32 + 127) / 128, (Fo + OUTPROJ_FC - 1) / OUTPROJ_FC), 128, 0, st>>>(n_node, H, Fo, h, w, b, out);
  GEOLDM_CHECK_LAUNCH("outproj_kernel");
  return 0;
}
__global__ void edge_dist_kernel(int n_edge, const int* __restrict__ ei, const int* __restrict__ ej,
                                 const float* __restrict__ x, float* __restrict__ out, float4* __restrict__ u_out,
                                 float norm_constant) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edge) return;
  const float* xi = x + 3 * (size_t)ei[e];
  const float* xj = x + 3 * (size_t)ej[e];
  const float dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
  const float r = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  out[e] = r;
  if (u_out) {      // same expression as edge_geom (common.cuh)
    const float den = __fadd_rn(__fsqrt_rn(__fadd_rn(r, 1e-8f)), norm_constant);
    u_out[e] = make_float4(__fdiv_rn(dx, den), __fdiv_rn(dy, den), __fdiv_rn(dz, den), 0.f);
  }
}
int launch_edge_dist(const geoldm_batch& b, const float* x, float* out, float* u_out, float norm_constant,
                     cudaStream_t st) {
  if (b.n_edge == 0) return 0;
  edge_dist_kernel<<<(b.n_edge + 255) / 256, 256, 0, st>>>(b.n_edge, b.edge_i, b.edge_j, x, out,
                                                           reinterpret_cast<float4*>(u_out), norm_constant);
  GEOLDM_CHECK_LAUNCH("edge_dist_kernel");
  return 0;
}
int launch_coord_update(int n3, const float* x0, const float* dx, float* xagg, float div, float* dx_next,
                        float* x_next, cudaStream_t st, bool zero_xagg) {
  if (n3 == 0) return 0;
  coord_update_kernel<<<(n3 + 255) / 256, 256, 0, st>>>(n3, x0, dx, xagg, div, dx_next, x_next, zero_xagg);
  GEOLDM_CHECK_LAUNCH("coord_update_kernel");
  return 0;
}
}  // namespace geoldm

"""CPU oracle for the GeoLDM EGNN-denoiser sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``geoldm_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker (or as the timed CPU baseline), never as the product path.

What it is: a functional (state-dict driven, no nn.Module) restatement, in plain torch CPU ops, of
the reference algorithm on the *padded* layout the reference uses.  It deliberately issues the
same kind of ATen ops the reference issues (gather, cat, Linear/addmm, scatter_add) so that its
CPU timing is representative of the reference's CPU path.  dtype-parametric: fp32 (parity target)
and fp64 (noise-floor reference).

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference from
/root/reference in the build container and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors (bit-level agreement is not
expected across BLAS builds; the tolerance is stated in the test).

Reference citations (paths relative to /root/reference):
  a1  coord2diff                      egnn/egnn_new.py:249-255
  a2  unsorted_segment_sum            egnn/egnn_new.py:258-274
  a3  GCL                             egnn/egnn_new.py:5-65
  a4  EquivariantUpdate               egnn/egnn_new.py:68-105
  a5  EquivariantBlock                egnn/egnn_new.py:108-147
  a6  EGNN                            egnn/egnn_new.py:150-197
  a7  EGNN_dynamics_QM9._forward      egnn/models.py:49-113 (+ get_adj_matrix :115-134)
  a8  remove_mean_with_mask           equivariant_diffusion/utils.py:31-38
  a9  combined position/feature noise equivariant_diffusion/en_diffusion.py:749-760, utils.py:107-116,137-140
  a10 PredefinedNoiseSchedule         equivariant_diffusion/en_diffusion.py:23-52,172-207
  a11 sigma/alpha t given s           equivariant_diffusion/en_diffusion.py:319-335,382-405
  a12 sample_p_zs_given_zt            equivariant_diffusion/en_diffusion.py:716-747
  a13 EnVariationalDiffusion.sample   equivariant_diffusion/en_diffusion.py:762-795
  a14 sample_p_xh_given_z0 (latent)   equivariant_diffusion/en_diffusion.py:1099-1122, 437-449
  a15 EnLatentDiffusion.sample/decode equivariant_diffusion/en_diffusion.py:1193-1204,1017-1035;
      EGNN_decoder_QM9._forward       egnn/models.py:335-381
  a16 qm9/sampling.py:sample          qm9/sampling.py:110-154
  a17 DistributionNodes               qm9/models.py:178-215
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Mirrors the argparse fields consumed by qm9/models.py:get_latent_diffusion (:103-166)."""
    nf: int = 256
    n_layers: int = 9
    inv_sublayers: int = 1
    latent_nf: int = 1
    attention: bool = True
    tanh: bool = True
    norm_constant: float = 1.0
    normalization_factor: float = 1.0
    aggregation_method: str = "sum"
    context_node_nf: int = 0
    condition_time: bool = True
    n_atom_types: int = 5            # len(dataset_info['atom_decoder'])
    include_charges: bool = True
    normalize_factors: Tuple[float, float, float] = (1.0, 4.0, 10.0)
    diffusion_steps: int = 1000
    diffusion_noise_schedule: str = "polynomial_2"
    diffusion_noise_precision: float = 1e-5
    coords_range: float = 15.0       # per block, NOT /n_layers
    n_dims: int = 3

    @property
    def data_node_nf(self) -> int:   # in_node_nf of the VAE (atom types + charge)
        return self.n_atom_types + int(self.include_charges)

    @property
    def dyn_in_nf(self) -> int:      # features entering the dynamics embedding
        return self.latent_nf + int(self.condition_time) + self.context_node_nf

    @property
    def dec_in_nf(self) -> int:
        return self.latent_nf + self.context_node_nf


QM9_CFG = OracleConfig()
GEOM_CFG = OracleConfig(nf=256, n_layers=4, latent_nf=2, n_atom_types=16, include_charges=False,
                        normalize_factors=(1.0, 4.0, 10.0))
QM9_COND_CFG = OracleConfig(nf=192, n_layers=9, latent_nf=1, context_node_nf=1, include_charges=False,
                            normalize_factors=(1.0, 8.0, 1.0))


# --------------------------------------------------------------------------------------------
# deterministic weights in the reference's state_dict layout
# --------------------------------------------------------------------------------------------
def egnn_param_shapes(in_nf: int, out_nf: int, H: int, L: int, S: int, attention: bool):
    """Ordered (key, shape, fan_in, kind) for one EGNN (egnn_new.py:150-182 registration order)."""
    out = [("embedding.weight", (H, in_nf), in_nf, "w"), ("embedding.bias", (H,), in_nf, "b"),
           ("embedding_out.weight", (out_nf, H), H, "w"), ("embedding_out.bias", (out_nf,), H, "b")]
    for b in range(L):
        for s in range(S):
            p = f"e_block_{b}.gcl_{s}."
            out += [(p + "edge_mlp.0.weight", (H, 2 * H + 2), 2 * H + 2, "w"), (p + "edge_mlp.0.bias", (H,), 2 * H + 2, "b"),
                    (p + "edge_mlp.2.weight", (H, H), H, "w"), (p + "edge_mlp.2.bias", (H,), H, "b"),
                    (p + "node_mlp.0.weight", (H, 2 * H), 2 * H, "w"), (p + "node_mlp.0.bias", (H,), 2 * H, "b"),
                    (p + "node_mlp.2.weight", (H, H), H, "w"), (p + "node_mlp.2.bias", (H,), H, "b")]
            if attention:
                out += [(p + "att_mlp.0.weight", (1, H), H, "w"), (p + "att_mlp.0.bias", (1,), H, "b")]
        p = f"e_block_{b}.gcl_equiv."
        out += [(p + "coord_mlp.0.weight", (H, 2 * H + 2), 2 * H + 2, "w"), (p + "coord_mlp.0.bias", (H,), 2 * H + 2, "b"),
                (p + "coord_mlp.2.weight", (H, H), H, "w"), (p + "coord_mlp.2.bias", (H,), H, "b"),
                (p + "coord_mlp.4.weight", (1, H), H, "x")]
    return out


def make_state_dict(cfg: OracleConfig, seed: int = 0, tamed: bool = False,
                    dtype=torch.float32, encoder: bool = False) -> StateDict:
    """Deterministic random-init weights (numpy PCG64, independent of torch's RNG/version).

    Same distributions as the PyTorch defaults the reference relies on: U(-1/sqrt(fan_in), +) for
    Linear weight and bias; coord_mlp.4 is xavier-uniform with gain 1e-3 (egnn_new.py:75-76).
    ``tamed`` applies SURVEY §8c's tamed init (distance columns x1e-5, dynamics embedding_out x0.01).
    Keys: 'dynamics.egnn.*', 'vae.decoder.egnn.*', 'gamma.gamma'; with ``encoder=True`` also
    'vae.encoder.egnn.*' (one block, egnn/models.py:153-160) and 'vae.encoder.final_mlp.{0,2}.*', drawn AFTER the
    others so that the dynamics/decoder weights of a seed do not depend on the flag.
    """
    rng = np.random.default_rng(seed)
    sd: StateDict = {}

    def fill(prefix, in_nf, out_nf, n_layers=cfg.n_layers):
        for key, shape, fan_in, kind in egnn_param_shapes(in_nf, out_nf, cfg.nf, n_layers,
                                                          cfg.inv_sublayers, cfg.attention):
            if kind == "x":
                bound = 1e-3 * math.sqrt(6.0 / (shape[0] + shape[1]))
            else:
                bound = 1.0 / math.sqrt(fan_in)
            arr = rng.uniform(-bound, bound, size=shape).astype(np.float32)
            sd[prefix + key] = torch.from_numpy(arr).to(dtype)

    fill("dynamics.egnn.", cfg.dyn_in_nf, cfg.dyn_in_nf)
    fill("vae.decoder.egnn.", cfg.dec_in_nf, cfg.data_node_nf)
    sd["gamma.gamma"] = torch.from_numpy(noise_schedule_gamma(cfg)).to(dtype)
    if encoder:
        H = cfg.nf
        fill("vae.encoder.egnn.", cfg.data_node_nf + cfg.context_node_nf, H, n_layers=1)
        for key, shape in (("final_mlp.0.weight", (H, H)), ("final_mlp.0.bias", (H,)),
                           ("final_mlp.2.weight", (2 * cfg.latent_nf + 1, H)),
                           ("final_mlp.2.bias", (2 * cfg.latent_nf + 1,))):
            bound = 1.0 / math.sqrt(H)
            sd["vae.encoder." + key] = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32)).to(dtype)
    if tamed:
        H = cfg.nf
        for k in list(sd):
            if k.endswith("edge_mlp.0.weight") or k.endswith("coord_mlp.0.weight"):
                w = sd[k].clone()
                w[:, 2 * H:] *= 1e-5
                sd[k] = w
        sd["dynamics.egnn.embedding_out.weight"] = sd["dynamics.egnn.embedding_out.weight"] * 0.01
        sd["dynamics.egnn.embedding_out.bias"] = sd["dynamics.egnn.embedding_out.bias"] * 0.01
    return sd


# --------------------------------------------------------------------------------------------
# a10: predefined polynomial noise schedule (float64 numpy, cast to fp32 like the reference)
# --------------------------------------------------------------------------------------------
def noise_schedule_gamma(cfg: OracleConfig) -> np.ndarray:
    """gamma[0..T] as float32 (en_diffusion.py:23-52 polynomial + clip, :172-203 table)."""
    name = cfg.diffusion_noise_schedule
    if not name.startswith("polynomial_"):
        raise ValueError(name)
    power = float(name.split("_")[1])
    T, s = cfg.diffusion_steps, cfg.diffusion_noise_precision
    steps = T + 1
    grid = np.linspace(0, steps, steps)
    a2 = (1.0 - np.power(grid / steps, power)) ** 2
    # clip alpha_t/alpha_{t-1} into [1e-3, 1]
    padded = np.concatenate([np.ones(1), a2])
    ratio = np.clip(padded[1:] / padded[:-1], 0.001, 1.0)
    a2 = np.cumprod(ratio)
    a2 = (1.0 - 2.0 * s) * a2 + s
    gamma = -(np.log(a2) - np.log(1.0 - a2))
    return gamma.astype(np.float32)


def gamma_lookup(gamma: Tensor, t: Tensor, T: int) -> Tensor:
    """PredefinedNoiseSchedule.forward (en_diffusion.py:205-207)."""
    return gamma[torch.round(t * T).long()]


# --------------------------------------------------------------------------------------------
# a1, a2
# --------------------------------------------------------------------------------------------
def coord2diff(x: Tensor, row: Tensor, col: Tensor, norm_constant: float):
    d = x.index_select(0, row) - x.index_select(0, col)
    radial = (d * d).sum(1, keepdim=True)
    d = d / (torch.sqrt(radial + 1e-8) + norm_constant)
    return radial, d


def segment_sum(data: Tensor, seg: Tensor, n_seg: int, normalization_factor: float, method: str) -> Tensor:
    out = torch.zeros(n_seg, data.shape[1], dtype=data.dtype)
    idx = seg.unsqueeze(1).expand(-1, data.shape[1])
    out.scatter_add_(0, idx, data)
    if method == "sum":
        out = out / normalization_factor
    elif method == "mean":                      # counts every edge in the segment, masked or not
        cnt = torch.zeros_like(out).scatter_add_(0, idx, torch.ones_like(data))
        cnt[cnt == 0] = 1
        out = out / cnt
    return out


def fully_connected_edges(bs: int, n: int) -> Tuple[Tensor, Tensor]:
    """All (i, j) pairs incl. self-edges, batch-major (egnn/models.py:115-134), vectorised."""
    base = (torch.arange(bs) * n).view(bs, 1, 1)
    i = torch.arange(n).view(1, n, 1)
    j = torch.arange(n).view(1, 1, n)
    row = (base + i + 0 * j).reshape(-1)
    col = (base + j + 0 * i).reshape(-1)
    return row, col


# --------------------------------------------------------------------------------------------
# a3..a6
# --------------------------------------------------------------------------------------------
def _lin(sd: StateDict, key: str, x: Tensor, bias: bool = True) -> Tensor:
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"] if bias else None)


def gcl(sd, p, cfg: OracleConfig, h, row, col, edge_attr, node_mask, edge_mask):
    e_in = torch.cat([h.index_select(0, row), h.index_select(0, col), edge_attr], dim=1)
    m = F.silu(_lin(sd, p + "edge_mlp.2", F.silu(_lin(sd, p + "edge_mlp.0", e_in))))
    if cfg.attention:
        m = m * torch.sigmoid(_lin(sd, p + "att_mlp.0", m))
    if edge_mask is not None:
        m = m * edge_mask
    agg = segment_sum(m, row, h.shape[0], cfg.normalization_factor, cfg.aggregation_method)
    upd = _lin(sd, p + "node_mlp.2", F.silu(_lin(sd, p + "node_mlp.0", torch.cat([h, agg], dim=1))))
    h = h + upd
    if node_mask is not None:
        h = h * node_mask
    return h


def equivariant_update(sd, p, cfg: OracleConfig, h, x, row, col, coord_diff, edge_attr, node_mask, edge_mask):
    e_in = torch.cat([h.index_select(0, row), h.index_select(0, col), edge_attr], dim=1)
    s = _lin(sd, p + "coord_mlp.4", F.silu(_lin(sd, p + "coord_mlp.2", F.silu(_lin(sd, p + "coord_mlp.0", e_in)))),
             bias=False)
    if cfg.tanh:
        trans = coord_diff * torch.tanh(s) * cfg.coords_range
    else:
        trans = coord_diff * s
    if edge_mask is not None:
        trans = trans * edge_mask
    x = x + segment_sum(trans, row, x.shape[0], cfg.normalization_factor, cfg.aggregation_method)
    if node_mask is not None:
        x = x * node_mask
    return x


def egnn(sd, prefix, cfg: OracleConfig, h, x, row, col, node_mask, edge_mask):
    d0, _ = coord2diff(x, row, col, 1.0)      # EGNN.forward calls coord2diff with the default constant
    h = _lin(sd, prefix + "embedding", h)
    for b in range(cfg.n_layers):
        bp = f"{prefix}e_block_{b}."
        r, u = coord2diff(x, row, col, cfg.norm_constant)
        ea = torch.cat([r, d0], dim=1)
        for s in range(cfg.inv_sublayers):
            h = gcl(sd, f"{bp}gcl_{s}.", cfg, h, row, col, ea, node_mask, edge_mask)
        x = equivariant_update(sd, bp + "gcl_equiv.", cfg, h, x, row, col, u, ea, node_mask, edge_mask)
        if node_mask is not None:
            h = h * node_mask
    h = _lin(sd, prefix + "embedding_out", h)
    if node_mask is not None:
        h = h * node_mask
    return h, x


# --------------------------------------------------------------------------------------------
# a8, a7, decoder
# --------------------------------------------------------------------------------------------
def remove_mean_with_mask(x: Tensor, node_mask: Tensor) -> Tensor:
    leak = (x * (1 - node_mask)).abs().sum().item()
    assert leak < 1e-5, f"Error {leak} too high"
    n = node_mask.sum(1, keepdim=True)
    return x - (x.sum(1, keepdim=True) / n) * node_mask


def assert_mean_zero_with_mask(x, node_mask, eps=1e-10):
    assert (x * (1 - node_mask)).abs().max().item() < 1e-4, "Variables not masked properly."
    largest = x.abs().max().item()
    err = x.sum(1, keepdim=True).abs().max().item()
    assert err / (largest + eps) < 1e-2, f"Mean is not zero, relative_error {err / (largest + eps)}"


def dynamics_forward(sd, cfg: OracleConfig, t, xh, node_mask, edge_mask, context=None,
                     prefix="dynamics.egnn.") -> Tensor:
    """EGNN_dynamics_QM9._forward: t scalar-tensor or [bs,1]; xh [bs,n,3+latent]."""
    bs, n, dims = xh.shape
    row, col = fully_connected_edges(bs, n)
    nm = node_mask.reshape(bs * n, 1)
    em = edge_mask.reshape(bs * n * n, 1)
    flat = xh.reshape(bs * n, dims) * nm
    x = flat[:, :cfg.n_dims].clone()
    h = flat[:, cfg.n_dims:].clone()
    if cfg.condition_time:
        if t.numel() == 1:
            h_time = torch.full_like(h[:, :1], float(t.reshape(-1)[0]))
        else:
            h_time = t.reshape(bs, 1).repeat(1, n).reshape(bs * n, 1).to(h.dtype)
        h = torch.cat([h, h_time], dim=1)
    if context is not None:
        h = torch.cat([h, context.reshape(bs * n, cfg.context_node_nf)], dim=1)
    h_f, x_f = egnn(sd, prefix, cfg, h, x, row, col, nm, em)
    vel = (x_f - x) * nm
    if context is not None:
        h_f = h_f[:, :-cfg.context_node_nf]
    if cfg.condition_time:
        h_f = h_f[:, :-1]
    vel = vel.reshape(bs, n, -1)
    if torch.isnan(vel).any():
        vel = torch.zeros_like(vel)
    vel = remove_mean_with_mask(vel, node_mask.reshape(bs, n, 1))
    return torch.cat([vel, h_f.reshape(bs, n, -1)], dim=2)


def decoder_forward(sd, cfg: OracleConfig, xh, node_mask, edge_mask, context=None,
                    prefix="vae.decoder.egnn."):
    """EGNN_decoder_QM9._forward: no time feature, x_out = x_final (not a difference)."""
    bs, n, dims = xh.shape
    row, col = fully_connected_edges(bs, n)
    nm = node_mask.reshape(bs * n, 1)
    em = edge_mask.reshape(bs * n * n, 1)
    flat = xh.reshape(bs * n, dims) * nm
    x = flat[:, :cfg.n_dims].clone()
    h = flat[:, cfg.n_dims:].clone()
    if context is not None:
        h = torch.cat([h, context.reshape(bs * n, cfg.context_node_nf)], dim=1)
    h_f, x_f = egnn(sd, prefix, cfg, h, x, row, col, nm, em)
    vel = (x_f * nm).reshape(bs, n, -1)
    if torch.isnan(vel).any():
        vel = torch.zeros_like(vel)
    vel = remove_mean_with_mask(vel, node_mask.reshape(bs, n, 1))
    h_f = (h_f * nm).reshape(bs, n, -1)
    return vel, h_f


# --------------------------------------------------------------------------------------------
# a9..a15: sampler
# --------------------------------------------------------------------------------------------
def _inflate(a: Tensor, target: Tensor) -> Tensor:
    return a.reshape((a.shape[0],) + (1,) * (target.dim() - 1))


def sigma_of(gamma, target):
    return _inflate(torch.sqrt(torch.sigmoid(gamma)), target)


def alpha_of(gamma, target):
    return _inflate(torch.sqrt(torch.sigmoid(-gamma)), target)


def sigma_and_alpha_t_given_s(gamma_t, gamma_s, target):
    sigma2 = _inflate(-torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t)), target)
    log_a2 = F.logsigmoid(-gamma_t) - F.logsigmoid(-gamma_s)
    alpha = _inflate(torch.exp(0.5 * log_a2), target)
    return sigma2, torch.sqrt(sigma2), alpha


def step_coefficients(gamma: Tensor, T: int, s: int):
    """Per-step scalars of a12 for integer s (t = s+1): returns python floats computed with the
    reference's fp32 formulae: (1/alpha_ts is NOT precomputed; the reference divides)."""
    g_s = gamma[s].reshape(1, 1)
    g_t = gamma[s + 1].reshape(1, 1)
    dummy = torch.zeros(1, 1, 1, dtype=gamma.dtype)
    sigma2_ts, sigma_ts, alpha_ts = sigma_and_alpha_t_given_s(g_t, g_s, dummy)
    sigma_s, sigma_t = sigma_of(g_s, dummy), sigma_of(g_t, dummy)
    c_eps = sigma2_ts / alpha_ts / sigma_t
    c_noise = sigma_ts * sigma_s / sigma_t
    return alpha_ts.reshape(()), c_eps.reshape(()), c_noise.reshape(())


class NoiseSource:
    """Noise provider. Default: torch.randn from the global generator in the reference's draw order
    (x block [bs,n,3] then h block [bs,n,latent], utils.py:107-116,137-140).  With ``raw`` set, pops
    pre-drawn float64 tensors RAW[k] of shape [bs,n,3+latent] in call order."""

    def __init__(self, raw: Optional[Tensor] = None):
        self.raw, self.k = raw, 0

    def draw(self, bs, n, n_dims, nf, dtype):
        if self.raw is None:
            zx = torch.randn(bs, n, n_dims, dtype=dtype)
            zh = torch.randn(bs, n, nf, dtype=dtype)
            return zx, zh
        r = self.raw[self.k].to(dtype)
        self.k += 1
        return r[..., :n_dims].clone(), r[..., n_dims:].clone()


def combined_noise(cfg: OracleConfig, noise: NoiseSource, bs, n, node_mask, nf) -> Tensor:
    zx, zh = noise.draw(bs, n, cfg.n_dims, nf, node_mask.dtype)
    zx = remove_mean_with_mask(zx * node_mask, node_mask)
    return torch.cat([zx, zh * node_mask], dim=2)


def sample_p_zs_given_zt(sd, cfg: OracleConfig, s: Tensor, t: Tensor, zt, node_mask, edge_mask, context,
                         noise: NoiseSource, return_eps: bool = False):
    gamma = sd["gamma.gamma"]
    T = cfg.diffusion_steps
    g_s, g_t = gamma_lookup(gamma, s, T), gamma_lookup(gamma, t, T)
    sigma2_ts, sigma_ts, alpha_ts = sigma_and_alpha_t_given_s(g_t, g_s, zt)
    sigma_s, sigma_t = sigma_of(g_s, zt), sigma_of(g_t, zt)
    eps_t = dynamics_forward(sd, cfg, t, zt, node_mask, edge_mask, context)
    assert_mean_zero_with_mask(zt[:, :, :cfg.n_dims], node_mask)
    assert_mean_zero_with_mask(eps_t[:, :, :cfg.n_dims], node_mask)
    mu = zt / alpha_ts - (sigma2_ts / alpha_ts / sigma_t) * eps_t
    sigma = sigma_ts * sigma_s / sigma_t
    eps = combined_noise(cfg, noise, mu.shape[0], mu.shape[1], node_mask, cfg.latent_nf)
    zs = mu + sigma * eps
    zs = torch.cat([remove_mean_with_mask(zs[:, :, :cfg.n_dims], node_mask), zs[:, :, cfg.n_dims:]], dim=2)
    return (zs, eps_t) if return_eps else zs


def sample_p_xh_given_z0(sd, cfg: OracleConfig, z0, node_mask, edge_mask, context, noise: NoiseSource):
    """EnLatentDiffusion override: returns (x, latent h) with no unnormalisation."""
    gamma = sd["gamma.gamma"]
    zeros = torch.zeros(z0.shape[0], 1, dtype=z0.dtype)
    g0 = gamma_lookup(gamma, zeros, cfg.diffusion_steps)
    sigma_x = torch.exp(0.5 * g0).unsqueeze(1)                  # SNR(-0.5 gamma_0)
    net = dynamics_forward(sd, cfg, zeros, z0, node_mask, edge_mask, context)
    mu = 1.0 / alpha_of(g0, net) * (z0 - sigma_of(g0, net) * net)
    eps = combined_noise(cfg, noise, mu.shape[0], mu.shape[1], node_mask, cfg.latent_nf)
    xh = mu + sigma_x * eps
    return xh[:, :, :cfg.n_dims], xh[:, :, cfg.n_dims:]


def sample_latent(sd, cfg: OracleConfig, bs, n, node_mask, edge_mask, context=None,
                  noise: Optional[NoiseSource] = None, trace=None, n_steps: Optional[int] = None):
    """EnVariationalDiffusion.sample: returns z_xh after p(x,h|z0).  ``trace`` (list) receives
    (z_t, eps_hat_t) per step for teacher-forced checks.  ``n_steps`` truncates the loop (testing)."""
    noise = noise or NoiseSource()
    T = cfg.diffusion_steps
    z = combined_noise(cfg, noise, bs, n, node_mask, cfg.latent_nf)
    assert_mean_zero_with_mask(z[:, :, :cfg.n_dims], node_mask)
    done = 0
    for s in reversed(range(T)):
        if n_steps is not None and done >= n_steps:
            return z
        s_arr = torch.full((bs, 1), float(s), dtype=z.dtype) / T
        t_arr = torch.full((bs, 1), float(s + 1), dtype=z.dtype) / T
        z_prev = z
        z, eps_t = sample_p_zs_given_zt(sd, cfg, s_arr, t_arr, z, node_mask, edge_mask, context, noise, True)
        if trace is not None:
            trace.append((z_prev, eps_t, z))
        done += 1
    x, h = sample_p_xh_given_z0(sd, cfg, z, node_mask, edge_mask, context, noise)
    assert_mean_zero_with_mask(x, node_mask)
    if x.sum(1, keepdim=True).abs().max().item() > 5e-2:
        x = remove_mean_with_mask(x, node_mask)
    return torch.cat([x, h], dim=2)


def decode(sd, cfg: OracleConfig, z_xh, node_mask, edge_mask, context=None):
    """EnHierarchicalVAE.decode incl. the h_cat = xh[:, :, 3:-1] quirk."""
    x_rec, h_rec = decoder_forward(sd, cfg, z_xh, node_mask, edge_mask, context)
    xh = torch.cat([x_rec, h_rec], dim=2)
    x = xh[:, :, :cfg.n_dims]
    num_classes = cfg.data_node_nf - int(cfg.include_charges)
    h_int = xh[:, :, -1:] if cfg.include_charges else torch.zeros(0, dtype=xh.dtype)
    h_cat = xh[:, :, cfg.n_dims:-1]
    one_hot = F.one_hot(torch.argmax(h_cat, dim=2), num_classes) * node_mask
    charges = torch.round(h_int).long() * node_mask
    return x, one_hot, charges


def build_masks(nodesxsample: Sequence[int], max_n_nodes: int, dtype=torch.float32):
    """node_mask [bs,n,1], edge_mask [bs*n*n,1] exactly as qm9/sampling.py:118-128."""
    n_arr = torch.as_tensor(list(map(int, nodesxsample)))
    nm = (torch.arange(max_n_nodes).unsqueeze(0) < n_arr.unsqueeze(1)).to(dtype)
    em = nm.unsqueeze(1) * nm.unsqueeze(2)
    em = em * (~torch.eye(max_n_nodes, dtype=torch.bool)).unsqueeze(0)
    return nm.unsqueeze(2), em.reshape(-1, 1)


def sample_molecules(sd, cfg: OracleConfig, nodesxsample, max_n_nodes, context=None,
                     noise: Optional[NoiseSource] = None, dtype=torch.float32):
    """qm9/sampling.py:sample -> (one_hot, charges, x, node_mask)."""
    node_mask, edge_mask = build_masks(nodesxsample, max_n_nodes, dtype)
    bs = node_mask.shape[0]
    if cfg.context_node_nf > 0:
        context = context.unsqueeze(1).repeat(1, max_n_nodes, 1) * node_mask
    else:
        context = None
    z = sample_latent(sd, cfg, bs, max_n_nodes, node_mask, edge_mask, context, noise)
    x, one_hot, charges = decode(sd, cfg, z, node_mask, edge_mask, context)
    return one_hot, charges, x, node_mask


# --------------------------------------------------------------------------------------------
# a17
# --------------------------------------------------------------------------------------------
def nodes_distribution_sample(histogram: Dict[int, int], n_samples: int) -> Tensor:
    """DistributionNodes.sample: categorical over the histogram keys in dict order (float64 probs),
    drawn with torch.distributions.Categorical from the global generator (qm9/models.py:178-201)."""
    keys = torch.tensor(list(histogram.keys()))
    prob = np.array(list(histogram.values()), dtype=np.float64)
    prob = prob / prob.sum()
    idx = torch.distributions.Categorical(torch.tensor(prob)).sample((n_samples,))
    return keys[idx]


def cast_state_dict(sd: StateDict, dtype) -> StateDict:
    return {k: v.to(dtype) for k, v in sd.items()}


def err_metric(a: Tensor, b: Tensor) -> float:
    """SURVEY §8c metric: max|a-b| / max|b|."""
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

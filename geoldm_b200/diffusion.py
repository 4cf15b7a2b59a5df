"""Sampling side of the reference's probabilistic model on the CUDA path.

Mirrors equivariant_diffusion/en_diffusion.py: ``PredefinedNoiseSchedule`` (:172-207),
``EnHierarchicalVAE.decode`` (:1017-1035) and ``EnLatentDiffusion`` (:1057-1232; ``sample`` :1193,
``sample_p_zs_given_zt`` :716, ``sample_p_xh_given_z0`` :1099, ``phi`` :314) with the same module tree, so
that a reference ``state_dict`` (``buffer``, ``gamma.gamma``, ``dynamics.egnn.*``, ``vae.*``) loads.

``EnLatentDiffusion.sample`` is the B200 hot loop: the latent state lives ragged-packed on the device,
every step is  prep -> EGNN kernels -> velocity/CoM -> z_s = mu + sigma*eps  with the step index and
per-step scalars in device memory, captured once as a CUDA graph and replayed T times; there is no host
synchronisation inside the loop (the reference does >= 10 per step).
The training objective (``forward``) lives in losses.py and runs the denoiser/decoder through train.py's autograd path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, losses
from .dynamics import EGNN_decoder_QM9, EGNN_dynamics_QM9, EGNN_encoder_QM9, _stream
from .packing import RaggedBatch, pack_from_masks


def polynomial_gamma(timesteps: int, precision: float, power: float) -> np.ndarray:
    """gamma[0..T] = -(log alpha^2 - log sigma^2) of the clipped polynomial schedule, float64
    (en_diffusion.py:23-52 and :192-203)."""
    steps = timesteps + 1
    u = np.linspace(0, steps, steps) / steps
    a2 = (1.0 - u ** power) ** 2
    ratio = np.clip(np.concatenate([a2[:1], a2[1:] / a2[:-1]]), 0.001, 1.0)   # first ratio is a2[0]/1
    a2 = (1.0 - 2.0 * precision) * np.cumprod(ratio) + precision
    return -(np.log(a2) - np.log(1.0 - a2))


class PredefinedNoiseSchedule(nn.Module):
    def __init__(self, noise_schedule, timesteps, precision):
        super().__init__()
        self.timesteps = timesteps
        if 'polynomial' not in noise_schedule:
            raise NotImplementedError("GeoLDM configs use polynomial_2; cosine/learned schedules are out of scope")
        parts = noise_schedule.split('_')
        assert len(parts) == 2
        gamma = polynomial_gamma(timesteps, precision, float(parts[1]))
        self.gamma = nn.Parameter(torch.from_numpy(gamma).float(), requires_grad=False)

    def forward(self, t):
        return self.gamma[torch.round(t * self.timesteps).long()]


class EnHierarchicalVAE(nn.Module):
    def __init__(self, encoder, decoder, in_node_nf: int, n_dims: int, latent_node_nf: int, kl_weight: float,
                 norm_values=(1., 1., 1.), norm_biases=(None, 0., 0.), include_charges=True):
        super().__init__()
        self.include_charges = include_charges
        self.encoder, self.decoder = encoder, decoder
        self.in_node_nf, self.n_dims, self.latent_node_nf = in_node_nf, n_dims, latent_node_nf
        self.num_classes = in_node_nf - int(include_charges)
        self.kl_weight = kl_weight
        self.norm_values, self.norm_biases = norm_values, norm_biases
        self.register_buffer('buffer', torch.zeros(1))

    def encode(self, x, h, node_mask=None, edge_mask=None, context=None):
        """q(z|x): encoder means and the fixed 0.0032 standard deviations (en_diffusion.py:1001-1015)."""
        xh = torch.cat([x, h['categorical'].to(x.dtype), h['integer'].to(x.dtype)], dim=2)
        z_x_mu, _, z_h_mu, _ = self.encoder._forward(xh, node_mask, edge_mask, context)
        bs = z_x_mu.shape[0]
        sigma_x = torch.full((bs, 1, 1), 0.0032, device=x.device, dtype=z_x_mu.dtype)
        sigma_h = torch.full((bs, 1, self.latent_node_nf), 0.0032, device=x.device, dtype=z_h_mu.dtype)
        return z_x_mu, sigma_x, z_h_mu, sigma_h

    def compute_reconstruction_error(self, xh_rec, xh):
        return losses.reconstruction_error(self, xh_rec, xh)

    def compute_loss(self, x, h, node_mask, edge_mask, context, *, draws=None):
        return losses.vae_loss(self, x, h, node_mask, edge_mask, context, draws=draws)

    def forward(self, x, h, node_mask=None, edge_mask=None, context=None, *, draws=None):
        """First-stage objective (en_diffusion.py:928-937)."""
        bs, n = x.shape[0], x.shape[1]
        edge_mask = None if edge_mask is None else edge_mask.reshape(bs * n * n, 1)
        return self.compute_loss(x, h, node_mask.reshape(bs, n, 1), edge_mask, context, draws=draws)[0]

    @torch.no_grad()
    def decode(self, z_xh, node_mask=None, edge_mask=None, context=None):
        """p(x|z): decoder EGNN, then argmax -> one_hot, round -> charges (on device, geoldm_decode).
        Reproduces the reference slice h_cat = xh[:, :, 3:-1] even when include_charges is False
        (en_diffusion.py:1030)."""
        x, h = self.decoder._forward(z_xh, node_mask, edge_mask, context)
        bs, n, Fo = h.shape
        h_flat = h.reshape(bs * n, Fo).contiguous()
        one_hot = torch.zeros(bs * n, self.num_classes, dtype=torch.int64, device=h.device)
        n_cat = Fo - 1
        charges = torch.zeros(bs * n, 1, dtype=torch.int64, device=h.device) if self.include_charges else None
        _lib.check(_lib.lib().geoldm_decode(bs * n, _lib.ptr(h_flat), Fo, n_cat, self.num_classes,
                                            Fo - 1 if self.include_charges else -1, _lib.ptr(one_hot),
                                            _lib.ptr(charges), _stream(h.device)), "geoldm_decode")
        nm = node_mask.reshape(bs, n, 1)
        one_hot = one_hot.view(bs, n, self.num_classes) * nm
        if self.include_charges:
            h_int = charges.view(bs, n, 1) * nm.long()
        else:
            h_int = torch.zeros(0, device=h.device) * nm   # reference: round(zeros(0)).long() * node_mask
        return x, {'integer': h_int, 'categorical': one_hot}


class EnLatentDiffusion(nn.Module):
    def __init__(self, vae: EnHierarchicalVAE, dynamics: EGNN_dynamics_QM9, in_node_nf: int, n_dims: int,
                 timesteps: int = 1000, parametrization='eps', noise_schedule='learned', noise_precision=1e-4,
                 loss_type='vlb', norm_values=(1., 1., 1.), norm_biases=(None, 0., 0.), include_charges=True,
                 trainable_ae=False):
        super().__init__()
        assert parametrization == 'eps'
        assert loss_type in {'vlb', 'l2'}
        self.loss_type, self.include_charges = loss_type, include_charges
        self.gamma = PredefinedNoiseSchedule(noise_schedule, timesteps=timesteps, precision=noise_precision)
        self.dynamics = dynamics
        self.in_node_nf, self.n_dims = in_node_nf, n_dims
        self.num_classes = in_node_nf - int(include_charges)
        self.T = timesteps
        self.parametrization = parametrization
        self.norm_values, self.norm_biases = norm_values, norm_biases
        self.register_buffer('buffer', torch.zeros(1))
        self.trainable_ae = trainable_ae
        self.vae = vae
        if not trainable_ae:
            self.vae.eval()
            for p in self.vae.parameters():
                p.requires_grad = False
        self.check_issues_norm_values()
        self._coef_cache = None
        self._graphs = {}
        self.use_cuda_graph = True

    def __getstate__(self):
        d = self.__dict__.copy()                 # coefficient table / captured graphs are rebuilt lazily
        d["_coef_cache"], d["_graphs"] = None, {}
        d.pop("_last_graph", None)
        return d

    def check_issues_norm_values(self, num_stdevs=8):
        sigma_0 = math.sqrt(1.0 / (1.0 + math.exp(-float(self.gamma.gamma[0]))))
        max_norm = max(self.norm_values[1], self.norm_values[2])
        if sigma_0 * num_stdevs > 1. / max_norm:
            raise ValueError(f'Value for normalization value {max_norm} probably too large with sigma_0 '
                             f'{sigma_0:.5f} and 1 / norm_value = {1. / max_norm}')

    # ---- per-step scalars (en_diffusion.py:327-335, 382-405, 733-736, 1103-1109) -----------------------
    def step_table(self, device) -> torch.Tensor:
        """[T+1, 4] fp32: rows s < T: {alpha_{t|s}, sigma2_{t|s}/alpha_{t|s}/sigma_t, sigma_{t|s} sigma_s / sigma_t,
        t=(s+1)/T}; row T: {1/alpha_0, sigma_0, exp(gamma_0/2), 0}.  Computed on the host with the reference's
        own fp32 tensor formulae so that the rounding of the scalars matches its CPU path."""
        key = (self.gamma.gamma.data_ptr(), self.gamma.gamma._version, str(device))
        if self._coef_cache is None or self._coef_cache[0] != key:
            g = self.gamma.gamma.detach().float().cpu()
            T = self.T
            g_s, g_t = g[:-1], g[1:]
            sigma2_ts = -torch.expm1(F.softplus(g_s) - F.softplus(g_t))
            alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
            sigma_ts = torch.sqrt(sigma2_ts)
            sigma_s, sigma_t = torch.sqrt(torch.sigmoid(g_s)), torch.sqrt(torch.sigmoid(g_t))
            t_val = (torch.arange(T) + 1) / T
            rows = torch.stack([alpha_ts, sigma2_ts / alpha_ts / sigma_t, sigma_ts * sigma_s / sigma_t, t_val], dim=1)
            g0 = g[:1]
            last = torch.stack([1.0 / torch.sqrt(torch.sigmoid(-g0)), torch.sqrt(torch.sigmoid(g0)),
                                torch.exp(0.5 * g0), torch.zeros(1)], dim=1)
            table = torch.cat([rows, last], dim=0).float().contiguous().to(device)
            self._coef_cache = (key, table)
        return self._coef_cache[1]

    def phi(self, x, t, node_mask, edge_mask, context):
        return self.dynamics._forward(t, x, node_mask, edge_mask, context)

    # ---- ragged fast path ---------------------------------------------------------------------------------
    def _denoise_ragged(self, batch: RaggedBatch, z, table, step_idx, ctx_r, eps_out):
        d = self.dynamics
        D = z.shape[1]
        d._run(batch, z, D, None, table, step_idx, ctx_r, d.condition_time, True, D - self.n_dims, eps_out, D,
               scatter=False)

    @torch.no_grad()
    def sample_latent_ragged(self, batch: RaggedBatch, context_ragged=None, fix_noise=False, noise=None,
                             seed: int = 0, n_steps: Optional[int] = None, trace: Optional[list] = None,
                             frame_cb=None):
        """a13+a14 on ragged state.  Returns z_xh [N, 3+latent] = (x, latent h) after p(x,h|z0).

        noise: None -> device Philox keyed by (seed, batch.mol_id) [fix_noise: every molecule uses key 0];
               tensor [T+2, N, D] (ragged, fp32, CUDA) -> injected draws in call order (init, steps, z0->x).
        n_steps: run only the first n_steps of the loop and return z (testing).
        trace: list receiving (z_t, eps_hat, z_s) clones per step (testing; disables the CUDA graph).
        frame_cb: callable(s, z) invoked on the host after the step that produced z_s (s = T-1 ... 0); it may enqueue
                  copies of z on the current stream (sample_chain)."""
        L = _lib.lib()
        dev = batch.mol_off.device
        D = self.n_dims + self.in_node_nf
        N, T = batch.n_node, self.T
        table = self.step_table(dev)
        cb = batch.c_batch(self.dynamics.egnn.tile_m())
        st = _stream(dev)
        z = torch.empty(N, D, device=dev)
        eps = torch.empty(N, D, device=dev)
        step_idx = torch.full((1,), T - 1, dtype=torch.int32, device=dev)
        draw_idx = torch.zeros(1, dtype=torch.int32, device=dev)
        mol_id = torch.zeros_like(batch.mol_id) if fix_noise else batch.mol_id
        nptr, nstride = (None, 0)
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.shape[1:] == (N, D), noise.shape
            noise = noise.contiguous()
            nptr, nstride = _lib.ptr(noise), N * D

        def update(mode, zin, zout):
            _lib.check(L.geoldm_sampler_update(C.byref(cb), mode, _lib.ptr(table), _lib.ptr(step_idx), _lib.ptr(zin),
                                               _lib.ptr(eps), nptr, nstride, D, C.c_uint64(seed), _lib.ptr(mol_id),
                                               _lib.ptr(draw_idx), _lib.ptr(zout), st), "geoldm_sampler_update")

        def advance(d_step, d_draw):
            _lib.check(L.geoldm_sampler_advance(_lib.ptr(step_idx), d_step, _lib.ptr(draw_idx), d_draw, st),
                       "geoldm_sampler_advance")

        def one_step():
            self._denoise_ragged(batch, z, table, step_idx, context_ragged, eps)
            update(0, z, z)
            advance(-1, 1)

        update(2, None, z)          # z_T ~ N(0, I) on the CoM-free subspace
        advance(0, 1)
        steps = T if n_steps is None else min(n_steps, T)
        if trace is not None or not self.use_cuda_graph or steps < 3:
            for k in range(steps):
                if trace is not None:
                    z_prev = z.clone()
                one_step()
                if trace is not None:
                    trace.append((z_prev, eps.clone(), z.clone()))
                if frame_cb is not None:
                    frame_cb(T - 1 - k, z)
        else:
            one_step()              # warm-up outside capture (lazy module init, workspace allocation)
            if frame_cb is not None:
                frame_cb(T - 1, z)
            graph = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cap):
                st = _stream(dev)
                with torch.cuda.graph(graph, stream=cap):
                    st = _stream(dev)
                    one_step()
            torch.cuda.current_stream(dev).wait_stream(cap)
            st = _stream(dev)
            for k in range(steps - 1):
                graph.replay()
                if frame_cb is not None:
                    frame_cb(T - 2 - k, z)
            self._last_graph = graph
        if n_steps is not None and steps < T:
            return z
        # p(x, h | z_0): one more denoiser call at t = 0 (table row T)
        advance(T + 1, 0)           # step_idx: -1 -> T
        self._denoise_ragged(batch, z, table, step_idx, context_ragged, eps)
        out = torch.empty_like(z)
        update(1, z, out)
        return out

    # ---- reference-shaped API -------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, n_samples, n_nodes, node_mask, edge_mask, context, fix_noise=False, *, noise=None, seed=0,
               mol_ids=None):
        """Draw samples (en_diffusion.py:1193-1204): returns (x [bs,n,3], {'categorical': one_hot, 'integer': charges}).

        Extra keyword-only arguments (not in the reference): ``noise`` injects pre-drawn normals
        [T+2, bs, n_nodes, 3+latent] (padded layout, reference call order) for parity runs; ``seed`` /
        ``mol_ids`` key the on-device Philox stream (per global molecule id, so results do not depend on how
        molecules are sharded over GPUs)."""
        if not node_mask.is_cuda:
            raise _lib.GeoldmError("geoldm_b200 has no CPU path: masks must be CUDA tensors")
        bs, n = n_samples, n_nodes
        node_mask = node_mask.reshape(bs, n, 1)
        batch = pack_from_masks(node_mask, edge_mask, validate=self.dynamics.validate_masks)
        if mol_ids is not None:
            batch.mol_id = torch.as_tensor(mol_ids, dtype=torch.int64, device=node_mask.device)
        src = batch.node_src.long()
        D = self.n_dims + self.in_node_nf
        ctx_r = None
        if context is not None:
            ctx_r = context.reshape(bs * n, -1)[src].contiguous().float()
        noise_r = None
        if noise is not None:
            noise_r = noise.to(node_mask.device, torch.float32).reshape(noise.shape[0], bs * n, D)[:, src].contiguous()
        z_r = self.sample_latent_ragged(batch, ctx_r, fix_noise=fix_noise, noise=noise_r, seed=seed)
        z_xh = torch.zeros(bs * n, D, device=node_mask.device)
        z_xh[src] = z_r
        z_xh = z_xh.view(bs, n, D)
        # cog-drift guard of EnVariationalDiffusion.sample (:789-793), evaluated once after the loop
        x = z_xh[:, :, :self.n_dims]
        max_cog = x.sum(dim=1, keepdim=True).abs().max()
        if bool(max_cog > 5e-2):
            nn_ = node_mask.sum(1, keepdim=True)
            x = x - (x.sum(1, keepdim=True) / nn_) * node_mask
            z_xh = torch.cat([x, z_xh[:, :, self.n_dims:]], dim=2)
        return self.vae.decode(z_xh, node_mask, edge_mask, context)

    @torch.no_grad()
    def sample_chain(self, n_samples, n_nodes, node_mask, edge_mask, context, keep_frames=None, *, noise=None, seed=0):
        """Sampling with intermediate states kept and decoded (en_diffusion.py:797-838 and :1206-1232): returns
        [keep_frames * n_samples, n_nodes, 3 + vae.in_node_nf]; frame 0 is the final sample, frame k the latent after the
        last step s with (s * keep_frames) // T == k.  Same loop and CUDA graph as ``sample``; frames are copied out
        between graph replays, then decoded one frame at a time like the reference."""
        if not node_mask.is_cuda:
            raise _lib.GeoldmError("geoldm_b200 has no CPU path: masks must be CUDA tensors")
        bs, n, T = n_samples, n_nodes, self.T
        kf = T if keep_frames is None else keep_frames
        assert kf <= T
        node_mask = node_mask.reshape(bs, n, 1)
        dev = node_mask.device
        batch = pack_from_masks(node_mask, edge_mask, validate=self.dynamics.validate_masks)
        src = batch.node_src.long()
        D = self.n_dims + self.in_node_nf
        ctx_r = None if context is None else context.reshape(bs * n, -1)[src].contiguous().float()
        noise_r = None
        if noise is not None:
            noise_r = noise.to(dev, torch.float32).reshape(noise.shape[0], bs * n, D)[:, src].contiguous()
        chain_r = torch.zeros(kf, batch.n_node, D, device=dev)

        def keep(s, z):
            idx = (s * kf) // T
            if s == 0 or ((s - 1) * kf) // T != idx:          # the reference overwrites; the smallest s of a frame wins
                chain_r[idx].copy_(z)

        chain_r[0] = self.sample_latent_ragged(batch, ctx_r, noise=noise_r, seed=seed, frame_cb=keep)
        out = torch.zeros(kf, bs, n, self.vae.in_node_nf + self.vae.n_dims, device=dev)
        for i in range(kf):
            z_xh = torch.zeros(bs * n, D, device=dev)
            z_xh[src] = chain_r[i]
            x, h = self.vae.decode(z_xh.view(bs, n, D), node_mask, edge_mask, context)
            out[i] = torch.cat([x, h['categorical'].to(x.dtype), h['integer'].to(x.dtype)], dim=2)
        return out.view(kf * bs, n, -1)

    @torch.no_grad()
    def sample_p_zs_given_zt(self, s, t, zt, node_mask, edge_mask, context, fix_noise=False, *, noise=None, seed=0,
                             draw=None):
        """One ancestral step on padded tensors (compatibility path; s, t must be uniform over the batch).
        Without injected ``noise`` the normals come from the Philox stream at draw index ``draw`` (default T - s, the
        numbering of ``sample_latent_ragged``: 0 = z_T, 1 = step T-1, ...), so a reference-style loop over s draws
        fresh noise at every step and reproduces ``sample`` with the same seed."""
        bs, n, D = zt.shape
        s_int = int(torch.round(s.reshape(-1)[0] * self.T).item())
        node_mask = node_mask.reshape(bs, n, 1)
        batch = self.dynamics._masks.get(node_mask, edge_mask, self.dynamics.validate_masks)
        src = batch.node_src.long()
        dev = zt.device
        eps = self.phi(zt, t, node_mask, edge_mask, context).reshape(bs * n, D)[src].contiguous()
        z_r = zt.reshape(bs * n, D)[src].contiguous()
        table = self.step_table(dev)
        step_idx = torch.full((1,), s_int, dtype=torch.int32, device=dev)
        noise_r = None if noise is None else noise.to(dev, torch.float32).reshape(bs * n, D)[src].contiguous()
        mol_id = torch.zeros_like(batch.mol_id) if fix_noise else batch.mol_id
        cb = batch.c_batch(self.dynamics.egnn.tile_m())
        out_r = torch.empty_like(z_r)
        draw_idx = torch.full((1,), self.T - s_int if draw is None else int(draw), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().geoldm_sampler_update(C.byref(cb), 0, _lib.ptr(table), _lib.ptr(step_idx), _lib.ptr(z_r),
                                                    _lib.ptr(eps), _lib.ptr(noise_r), 0, D, C.c_uint64(seed),
                                                    _lib.ptr(mol_id), _lib.ptr(draw_idx), _lib.ptr(out_r), _stream(dev)),
                   "geoldm_sampler_update")
        out = torch.zeros(bs * n, D, device=dev)
        out[src] = out_r
        return out.view(bs, n, D)

    def train(self, mode: bool = True):
        """A frozen first stage stays in eval() whatever the outer mode is (en_diffusion.py:1234-1239)."""
        super().train(mode)
        if not self.trainable_ae:
            self.vae.eval()
        return self

    def forward(self, x, h, node_mask=None, edge_mask=None, context=None, *, draws=None):
        """Per-molecule training loss (train(): l2) or NLL estimate (eval()), en_diffusion.py:1136-1191.
        ``draws`` (keyword-only, not in the reference) injects the random draws for parity tests."""
        bs, n = x.shape[0], x.shape[1]
        edge_mask = None if edge_mask is None else edge_mask.reshape(bs * n * n, 1)
        return losses.latent_diffusion_nll(self, x, h, node_mask.reshape(bs, n, 1), edge_mask, context, draws=draws)

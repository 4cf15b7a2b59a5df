"""Drop-in replacements for the reference's network wrappers (egnn/models.py) on the CUDA path.

``EGNN_dynamics_QM9`` keeps the reference constructor (egnn/models.py:9-13), attribute names and the
``_forward(t, xh, node_mask, edge_mask, context)`` contract (:49-113): padded ``[bs, n, 3+latent]``
in, padded ``[bs, n, 3+latent]`` out (masked rows exactly zero).  Internally molecules are packed
ragged, and every arithmetic step runs in the hand-written kernels behind the C ABI.
Differences a maintainer should know:
  * no host synchronisation: the NaN guard (models.py:100-102) is applied on device and the flag is
    left in ``self.nan_flag`` instead of printing; mask assertions are validated once per new mask;
  * ``mode='gnn_dynamics'`` and ``sin_embedding`` are not used by any GeoLDM config -> NotImplementedError.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib
from .egnn import EGNN
from .packing import RaggedBatch, pack_from_masks
from .train import linear as train_linear, wrapper_forward_train


class _MaskCache:
    """node_mask -> RaggedBatch, keyed by storage identity (the sampler passes the same mask 1000x)."""

    def __init__(self):
        self.key, self.batch, self.hold = None, None, None

    def get(self, node_mask, edge_mask, validate=True) -> RaggedBatch:
        key = (node_mask.data_ptr(), tuple(node_mask.shape), node_mask._version,
               None if edge_mask is None else (edge_mask.data_ptr(), edge_mask._version))
        if key != self.key:
            self.batch = pack_from_masks(node_mask, edge_mask, validate=validate)
            self.key = key
            # strong refs: while they are alive their storage (hence data_ptr) cannot be handed to another mask
            self.hold = (node_mask, edge_mask)
        return self.batch


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _EgnnWrapper(nn.Module):
    def __getstate__(self):
        d = self.__dict__.copy()                 # caches (packed masks, scratch) are per-instance and rebuilt lazily
        d["_masks"] = _MaskCache()
        d.pop("_bufs", None)
        return d

    def get_adj_matrix(self, n_nodes, batch_size, device):
        """Reference API (egnn/models.py:115-134).  The CUDA path never materialises the edge list;
        provided (vectorised) for callers that want it."""
        base = (torch.arange(batch_size, device=device) * n_nodes).view(-1, 1, 1)
        i = torch.arange(n_nodes, device=device).view(1, -1, 1)
        j = torch.arange(n_nodes, device=device).view(1, 1, -1)
        return [(base + i + 0 * j).reshape(-1), (base + j + 0 * i).reshape(-1)]

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def wrap_forward(self, node_mask, edge_mask, context):
        def fwd(time, state):
            return self._forward(time, state, node_mask, edge_mask, context)
        return fwd

    def unwrap_forward(self):
        return self._forward

    # ragged core shared by dynamics / decoder -------------------------------------------------------
    def _run(self, batch: RaggedBatch, xh_flat, xh_dim, t_mol, t_table, step_idx, ctx_flat, condition_time, delta,
             h_keep, out, out_dim, scatter: bool):
        """prep -> EGNN -> velocity / NaN guard / CoM -> out.  All tensors fp32 CUDA; no sync."""
        L = _lib.lib()
        dev = xh_flat.device
        egnn = self.egnn
        N = batch.n_node
        cb = batch.c_batch(egnn.tile_m())
        st = _stream(dev)
        buf = self._scratch(N, dev)
        src = batch.node_src if scatter else None
        _lib.check(L.geoldm_dynamics_prep(C.byref(cb), _lib.ptr(src), _lib.ptr(xh_flat), xh_dim, _lib.ptr(t_mol),
                                          _lib.ptr(t_table), _lib.ptr(step_idx), _lib.ptr(ctx_flat),
                                          self.context_node_nf if ctx_flat is not None else 0, int(condition_time),
                                          _lib.ptr(buf["h_in"]), egnn.in_node_nf, _lib.ptr(buf["x"]), st),
                   "geoldm_dynamics_prep")
        egnn.forward(buf["h_in"], buf["x"], batch, h_out=buf["h_out"], x_out=buf["x_out"], dx_out=buf["vel"])
        vel = buf["vel"] if delta else buf["x_out"]      # dynamics: x_final - x ; decoder: x_final
        self.nan_flag.zero_()
        _lib.check(L.geoldm_dynamics_finish_a(C.byref(cb), _lib.ptr(vel), _lib.ptr(self.nan_flag), st),
                   "geoldm_dynamics_finish_a")
        _lib.check(L.geoldm_dynamics_finish_b(C.byref(cb), _lib.ptr(src), _lib.ptr(vel), _lib.ptr(buf["h_out"]),
                                              egnn.out_node_nf, h_keep, _lib.ptr(self.nan_flag), _lib.ptr(out),
                                              out_dim, st), "geoldm_dynamics_finish_b")
        return out

    def _scratch(self, N, dev):
        b = getattr(self, "_bufs", None)
        if b is None or b["N"] != N or b["x"].device != dev:
            e = self.egnn
            b = dict(N=N, h_in=torch.empty(N, e.in_node_nf, device=dev), x=torch.empty(N, 3, device=dev),
                     h_out=torch.empty(N, e.out_node_nf, device=dev), x_out=torch.empty(N, 3, device=dev),
                     vel=torch.empty(N, 3, device=dev))
            self._bufs = b
        return b

    def _wants_grad(self, xh, context):
        """Autograd path (train.py) whenever a gradient can be asked for, like the reference, which always builds the
        graph: grad mode on and (train() mode, or an input requires grad, or any EGNN parameter requires grad — e.g.
        eval()-mode NLL fine-tuning).  Sampling and evaluation run under torch.no_grad() (en_diffusion.py:762,1193) and
        take the fused inference kernels."""
        if not torch.is_grad_enabled():
            return False
        if self.training or xh.requires_grad or (context is not None and context.requires_grad):
            return True
        return any(p.requires_grad for p in self.egnn.parameters())

    def _check_inputs(self, xh, node_mask):
        if not xh.is_cuda:
            raise _lib.GeoldmError("geoldm_b200 has no CPU path: inputs must be CUDA tensors")
        if xh.dtype != torch.float32:
            raise TypeError("the CUDA path computes in fp32; got " + str(xh.dtype))


class EGNN_dynamics_QM9(_EgnnWrapper):
    def __init__(self, in_node_nf, context_node_nf, n_dims, hidden_nf=64, device='cpu', act_fn=torch.nn.SiLU(),
                 n_layers=4, attention=False, condition_time=True, tanh=False, mode='egnn_dynamics', norm_constant=0,
                 inv_sublayers=2, sin_embedding=False, normalization_factor=100, aggregation_method='sum',
                 mma_mode="fp32", validate_masks=True):
        super().__init__()
        if mode != 'egnn_dynamics':
            raise NotImplementedError("only mode='egnn_dynamics' is on the GeoLDM hot path (gnn_dynamics unsupported)")
        if n_dims != 3:
            raise NotImplementedError("n_dims must be 3")
        self.mode = mode
        self.egnn = EGNN(in_node_nf=in_node_nf + context_node_nf, in_edge_nf=1, hidden_nf=hidden_nf, device=device,
                         act_fn=act_fn, n_layers=n_layers, attention=attention, tanh=tanh, norm_constant=norm_constant,
                         inv_sublayers=inv_sublayers, sin_embedding=sin_embedding,
                         normalization_factor=normalization_factor, aggregation_method=aggregation_method,
                         mma_mode=mma_mode)
        self.in_node_nf = in_node_nf
        self.context_node_nf = context_node_nf
        self.device = device
        self.n_dims = n_dims
        self._edges_dict = {}
        self.condition_time = condition_time
        self.validate_masks = validate_masks
        self._masks = _MaskCache()
        self.register_buffer("nan_flag", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.to(device)

    def _forward(self, t, xh, node_mask, edge_mask, context):
        self._check_inputs(xh, node_mask)
        if self._wants_grad(xh, context):
            return wrapper_forward_train(self, t, xh, node_mask, edge_mask, context, True)
        with torch.no_grad():
            return self._forward_infer(t, xh, node_mask, edge_mask, context)

    def _forward_infer(self, t, xh, node_mask, edge_mask, context):
        bs, n_nodes, dims = xh.shape
        batch = self._masks.get(node_mask.view(bs, n_nodes, 1), edge_mask, self.validate_masks)
        h_dims = dims - self.n_dims
        if h_dims == 0:
            raise NotImplementedError("h_dims == 0 is not used by GeoLDM (latent_nf >= 1)")
        xh_flat = xh.reshape(bs * n_nodes, dims).contiguous()
        t_mol = None
        if self.condition_time:
            t = torch.as_tensor(t, dtype=torch.float32, device=xh.device)
            t_mol = (t.reshape(1).expand(bs) if t.numel() == 1 else t.reshape(bs)).contiguous()
        ctx = None
        if context is not None:
            ctx = context.reshape(bs * n_nodes, self.context_node_nf).contiguous().float()
        keep = self.egnn.out_node_nf - self.context_node_nf - int(self.condition_time)
        out = torch.zeros(bs * n_nodes, self.n_dims + keep, device=xh.device)
        self._run(batch, xh_flat, dims, t_mol, None, None, ctx, self.condition_time, True, keep, out,
                  self.n_dims + keep, scatter=True)
        return out.view(bs, n_nodes, self.n_dims + keep)


class EGNN_decoder_QM9(_EgnnWrapper):
    """egnn/models.py:287-381: same EGNN, no time feature, x_out = x_final."""

    def __init__(self, in_node_nf, context_node_nf, out_node_nf, n_dims, hidden_nf=64, device='cpu',
                 act_fn=torch.nn.SiLU(), n_layers=4, attention=False, tanh=False, mode='egnn_dynamics',
                 norm_constant=0, inv_sublayers=2, sin_embedding=False, normalization_factor=100,
                 aggregation_method='sum', include_charges=True, mma_mode="fp32", validate_masks=True):
        super().__init__()
        if mode != 'egnn_dynamics':
            raise NotImplementedError("only mode='egnn_dynamics' is supported")
        include_charges = int(include_charges)
        self.mode = mode
        self.egnn = EGNN(in_node_nf=in_node_nf + context_node_nf, out_node_nf=out_node_nf, in_edge_nf=1,
                         hidden_nf=hidden_nf, device=device, act_fn=act_fn, n_layers=n_layers, attention=attention,
                         tanh=tanh, norm_constant=norm_constant, inv_sublayers=inv_sublayers,
                         sin_embedding=sin_embedding, normalization_factor=normalization_factor,
                         aggregation_method=aggregation_method, mma_mode=mma_mode)
        self.in_node_nf = in_node_nf
        self.num_classes = out_node_nf - include_charges
        self.include_charges = include_charges
        self.context_node_nf = context_node_nf
        self.device = device
        self.n_dims = n_dims
        self._edges_dict = {}
        self.validate_masks = validate_masks
        self._masks = _MaskCache()
        self.register_buffer("nan_flag", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.to(device)

    def _forward(self, xh, node_mask, edge_mask, context):
        self._check_inputs(xh, node_mask)
        if self._wants_grad(xh, context):
            return wrapper_forward_train(self, None, xh, node_mask, edge_mask, context, False)
        with torch.no_grad():
            return self._forward_infer(xh, node_mask, edge_mask, context)

    def _forward_infer(self, xh, node_mask, edge_mask, context):
        bs, n_nodes, dims = xh.shape
        batch = self._masks.get(node_mask.view(bs, n_nodes, 1), edge_mask, self.validate_masks)
        xh_flat = xh.reshape(bs * n_nodes, dims).contiguous()
        ctx = None
        if context is not None:
            ctx = context.reshape(bs * n_nodes, self.context_node_nf).contiguous().float()
        Fo = self.egnn.out_node_nf
        out = torch.zeros(bs * n_nodes, self.n_dims + Fo, device=xh.device)
        self._run(batch, xh_flat, dims, None, None, None, ctx, False, False, Fo, out, self.n_dims + Fo, scatter=True)
        out = out.view(bs, n_nodes, self.n_dims + Fo)
        return out[:, :, :self.n_dims], out[:, :, self.n_dims:]


class EGNN_encoder_QM9(_EgnnWrapper):
    """egnn/models.py:137-263: one-block EGNN (out = hidden_nf) + final_mlp -> (x mean, x std, h mean, h std).
    Inside EnLatentDiffusion the encoder is always evaluated without grad (its output is detached,
    en_diffusion.py:1155) and runs on the fused inference kernels; with grad (first-stage training) it goes through
    train.py like the other wrappers."""

    def __init__(self, in_node_nf, context_node_nf, out_node_nf, n_dims, hidden_nf=64, device='cpu',
                 act_fn=torch.nn.SiLU(), n_layers=4, attention=False, tanh=False, mode='egnn_dynamics',
                 norm_constant=0, inv_sublayers=2, sin_embedding=False, normalization_factor=100,
                 aggregation_method='sum', include_charges=True, mma_mode="fp32", validate_masks=True):
        super().__init__()
        if mode != 'egnn_dynamics':
            raise NotImplementedError("only mode='egnn_dynamics' is supported")
        self.mode = mode
        self.egnn = EGNN(in_node_nf=in_node_nf + context_node_nf, out_node_nf=hidden_nf, in_edge_nf=1,
                         hidden_nf=hidden_nf, device=device, act_fn=act_fn, n_layers=n_layers, attention=attention,
                         tanh=tanh, norm_constant=norm_constant, inv_sublayers=inv_sublayers,
                         sin_embedding=sin_embedding, normalization_factor=normalization_factor,
                         aggregation_method=aggregation_method, mma_mode=mma_mode)
        self.final_mlp = nn.Sequential(nn.Linear(hidden_nf, hidden_nf), nn.SiLU(),
                                       nn.Linear(hidden_nf, out_node_nf * 2 + 1))
        self.in_node_nf, self.out_node_nf = in_node_nf, out_node_nf
        self.num_classes = in_node_nf - int(include_charges)
        self.include_charges = int(include_charges)
        self.context_node_nf, self.n_dims, self.device = context_node_nf, n_dims, device
        self._edges_dict = {}
        self.validate_masks = validate_masks
        self._masks = _MaskCache()
        self.register_buffer("nan_flag", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.to(device)

    def _forward(self, xh, node_mask, edge_mask, context):
        self._check_inputs(xh, node_mask)
        bs, n_nodes, dims = xh.shape
        nm = node_mask.reshape(bs, n_nodes, 1)
        if self._wants_grad(xh, context):
            vel, h_final = wrapper_forward_train(self, None, xh, node_mask, edge_mask, context, False)
            h_final = train_linear(torch.nn.functional.silu(train_linear(
                h_final.reshape(bs * n_nodes, -1), self.final_mlp[0].weight, self.final_mlp[0].bias)),
                self.final_mlp[2].weight, self.final_mlp[2].bias).view(bs, n_nodes, -1) * nm
        else:
            with torch.no_grad():
                batch = self._masks.get(nm, edge_mask, self.validate_masks)
                xh_flat = xh.reshape(bs * n_nodes, dims).contiguous()
                ctx = None
                if context is not None:
                    ctx = context.reshape(bs * n_nodes, self.context_node_nf).contiguous().float()
                Fo = self.egnn.out_node_nf
                out = torch.zeros(bs * n_nodes, self.n_dims + Fo, device=xh.device)
                self._run(batch, xh_flat, dims, None, None, None, ctx, False, False, Fo, out, self.n_dims + Fo,
                          scatter=True)
                vel = out[:, :self.n_dims].reshape(bs, n_nodes, self.n_dims)
                h_final = self.final_mlp(out[:, self.n_dims:]).view(bs, n_nodes, -1) * nm
        vel_std = torch.exp(0.5 * h_final[:, :, :1].sum(dim=1, keepdim=True).expand(-1, n_nodes, -1))
        h_mean = h_final[:, :, 1:1 + self.out_node_nf]
        h_std = torch.exp(0.5 * h_final[:, :, 1 + self.out_node_nf:])
        vel_std = torch.where(torch.isnan(vel_std).any(), torch.ones_like(vel_std), vel_std)
        h_std = torch.where(torch.isnan(h_std).any(), torch.ones_like(h_std), h_std)
        return vel, vel_std, h_mean, h_std

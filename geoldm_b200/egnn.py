"""Host-side mirror of the reference's EGNN modules (egnn/egnn_new.py) for the CUDA path.

The classes keep the reference's constructor arguments and — because checkpoints are plain
``state_dict`` pickles (utils.py:20-27) — its parameter names and shapes:
``embedding``, ``embedding_out``, ``e_block_{b}.gcl_{s}.{edge_mlp.0,edge_mlp.2,node_mlp.0,node_mlp.2,att_mlp.0}``,
``e_block_{b}.gcl_equiv.coord_mlp.{0,2,4}``.  No arithmetic happens in Python: ``EGNN.forward`` packs
the weights once (first-layer split into per-node projections, transposes) and calls
``geoldm_egnn_forward`` through the C ABI on ragged-packed inputs.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch import nn

from . import _lib
from .packing import RaggedBatch

TILE_M = {_lib.MMA_FP32_SIMT: 64, _lib.MMA_3XTF32: 128, _lib.MMA_TF32: 128, _lib.MMA_BF16: 128,
          _lib.MMA_3XF16: 128}


def _require_silu(act_fn):
    if not isinstance(act_fn, nn.SiLU):
        raise NotImplementedError("the fused kernels implement SiLU only (all GeoLDM configs use SiLU)")


def _mlp(*dims_and_acts):
    return nn.Sequential(*dims_and_acts)


class GCL(nn.Module):
    """Parameter container of one graph-convolution layer (egnn_new.py:5-28)."""

    def __init__(self, input_nf, output_nf, hidden_nf, normalization_factor, aggregation_method,
                 edges_in_d=0, nodes_att_dim=0, act_fn=nn.SiLU(), attention=False):
        super().__init__()
        _require_silu(act_fn)
        if input_nf != hidden_nf or output_nf != hidden_nf or nodes_att_dim != 0 or edges_in_d != 2:
            raise NotImplementedError("fused GCL kernels need input_nf == output_nf == hidden_nf, 2 edge features")
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        self.attention = attention
        self.edge_mlp = _mlp(nn.Linear(2 * input_nf + edges_in_d, hidden_nf), nn.SiLU(),
                             nn.Linear(hidden_nf, hidden_nf), nn.SiLU())
        self.node_mlp = _mlp(nn.Linear(hidden_nf + input_nf, hidden_nf), nn.SiLU(), nn.Linear(hidden_nf, output_nf))
        if attention:
            self.att_mlp = _mlp(nn.Linear(hidden_nf, 1), nn.Sigmoid())


class EquivariantUpdate(nn.Module):
    """Parameter container of the coordinate update (egnn_new.py:68-84)."""

    def __init__(self, hidden_nf, normalization_factor, aggregation_method, edges_in_d=1, act_fn=nn.SiLU(),
                 tanh=False, coords_range=10.0):
        super().__init__()
        _require_silu(act_fn)
        self.tanh = tanh
        self.coords_range = coords_range
        head = nn.Linear(hidden_nf, 1, bias=False)
        nn.init.xavier_uniform_(head.weight, gain=0.001)
        self.coord_mlp = _mlp(nn.Linear(2 * hidden_nf + edges_in_d, hidden_nf), nn.SiLU(),
                              nn.Linear(hidden_nf, hidden_nf), nn.SiLU(), head)
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method


class EquivariantBlock(nn.Module):
    """S GCLs followed by one EquivariantUpdate (egnn_new.py:108-132)."""

    def __init__(self, hidden_nf, edge_feat_nf=2, device='cpu', act_fn=nn.SiLU(), n_layers=2, attention=True,
                 norm_diff=True, tanh=False, coords_range=15, norm_constant=1, sin_embedding=None,
                 normalization_factor=100, aggregation_method='sum'):
        super().__init__()
        if sin_embedding is not None:
            raise NotImplementedError("sin_embedding is not used by any GeoLDM config and is not implemented")
        self.hidden_nf, self.device, self.n_layers = hidden_nf, device, n_layers
        self.coords_range_layer = float(coords_range)
        self.norm_diff, self.norm_constant = norm_diff, norm_constant
        self.normalization_factor, self.aggregation_method = normalization_factor, aggregation_method
        for i in range(n_layers):
            self.add_module(f"gcl_{i}", GCL(hidden_nf, hidden_nf, hidden_nf, edges_in_d=edge_feat_nf, act_fn=act_fn,
                                            attention=attention, normalization_factor=normalization_factor,
                                            aggregation_method=aggregation_method))
        self.add_module("gcl_equiv", EquivariantUpdate(hidden_nf, edges_in_d=edge_feat_nf, act_fn=nn.SiLU(), tanh=tanh,
                                                       coords_range=self.coords_range_layer,
                                                       normalization_factor=normalization_factor,
                                                       aggregation_method=aggregation_method))
        self.to(device)


class EGNN(nn.Module):
    """egnn_new.py:150-197 on the CUDA path.  ``forward`` takes ragged-packed tensors."""

    def __init__(self, in_node_nf, in_edge_nf, hidden_nf, device='cpu', act_fn=nn.SiLU(), n_layers=3, attention=False,
                 norm_diff=True, out_node_nf=None, tanh=False, coords_range=15, norm_constant=1, inv_sublayers=2,
                 sin_embedding=False, normalization_factor=100, aggregation_method='sum', mma_mode="fp32"):
        super().__init__()
        _require_silu(act_fn)
        if sin_embedding:
            raise NotImplementedError("sin_embedding is not used by any GeoLDM config and is not implemented")
        if aggregation_method not in ("sum", "mean"):
            raise ValueError(aggregation_method)
        if n_layers > _lib.MAX_LAYERS or inv_sublayers > _lib.MAX_SUBLAYERS:
            raise ValueError(f"at most {_lib.MAX_LAYERS} blocks x {_lib.MAX_SUBLAYERS} sublayers")
        self.in_node_nf = in_node_nf
        self.out_node_nf = in_node_nf if out_node_nf is None else out_node_nf
        self.hidden_nf, self.device, self.n_layers, self.inv_sublayers = hidden_nf, device, n_layers, inv_sublayers
        # kept for parity with the reference attribute (unused there as well: each block gets coords_range)
        self.coords_range_layer = float(coords_range / n_layers) if n_layers > 0 else float(coords_range)
        self.coords_range = float(coords_range)
        self.norm_diff, self.norm_constant = norm_diff, norm_constant
        self.normalization_factor, self.aggregation_method = normalization_factor, aggregation_method
        self.attention, self.tanh = attention, tanh
        self.sin_embedding = None
        if mma_mode == "auto":      # fastest parity-green arithmetic this hidden size supports (tensor-core tiles need
            mma_mode = "3xf16" if (hidden_nf % 64 == 0 and hidden_nf <= 256) else "fp32"   # H in {64, 128, 192, 256})
        self.mma_mode = mma_mode
        self.embedding = nn.Linear(in_node_nf, hidden_nf)
        self.embedding_out = nn.Linear(hidden_nf, self.out_node_nf)
        for i in range(n_layers):
            self.add_module(f"e_block_{i}", EquivariantBlock(
                hidden_nf, edge_feat_nf=2, device=device, act_fn=act_fn, n_layers=inv_sublayers, attention=attention,
                norm_diff=norm_diff, tanh=tanh, coords_range=coords_range, norm_constant=norm_constant,
                sin_embedding=None, normalization_factor=normalization_factor, aggregation_method=aggregation_method))
        self.to(device)
        self._pack = None
        self._pack_key = None
        self._ws = None

    # ---- weight packing -------------------------------------------------------------------------
    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def __getstate__(self):
        """copy.deepcopy / pickle (the reference clones the model for its EMA copy, main_qm9.py:227-231): derived
        device images and scratch are rebuilt on demand, never copied."""
        d = self.__dict__.copy()
        d["_pack"], d["_pack_key"], d["_ws"] = None, None, None
        return d

    @torch.no_grad()
    def packed(self):
        """(ctypes EgnnWeights, keep-alive list).  Rebuilt when any parameter changed."""
        key = (self._params_key(), self.mma_mode)
        if self._pack is not None and key == self._pack_key:
            return self._pack
        H = self.hidden_nf
        keep = []

        def dev(t):
            t = t.detach().to(torch.float32).contiguous()
            if not t.is_cuda:
                raise _lib.GeoldmError("EGNN parameters must live on a CUDA device (no CPU path)")
            keep.append(t)
            return t.data_ptr()

        mode = _lib.MMA_MODES[self.mma_mode] if isinstance(self.mma_mode, str) else int(self.mma_mode)
        tcore = mode != _lib.MMA_FP32_SIMT
        L = _lib.lib()
        stream = C.c_void_p(torch.cuda.current_stream(self.embedding.weight.device).cuda_stream) \
            if self.embedding.weight.is_cuda else None

        def tc_pack(wmat):
            """[n_out, k] Linear weight -> tensor-core operand pack (tf32 hi/lo, SWIZZLE_128B k-slabs)."""
            if not tcore:
                return None
            wmat = wmat.detach().to(torch.float32).contiguous()
            n_out, k = wmat.shape
            if mode == _lib.MMA_3XF16:
                buf = torch.empty(L.geoldm_tc_pack16_bytes(H, n_out, k), dtype=torch.uint8, device=wmat.device)
                _lib.check(L.geoldm_tc_pack16(H, _lib.ptr(wmat), n_out, k, _lib.ptr(buf), stream), "geoldm_tc_pack16")
            else:
                buf = torch.empty(L.geoldm_tc_pack_bytes(H, n_out, k), dtype=torch.uint8, device=wmat.device)
                _lib.check(L.geoldm_tc_pack(H, _lib.ptr(wmat), n_out, k, _lib.ptr(buf), stream), "geoldm_tc_pack")
            keep.append(wmat)
            keep.append(buf)
            return buf.data_ptr()

        def edge(first, second, head, head_bias):
            w1 = first.weight                                    # [H, 2H+2]
            e = _lib.EdgeMlp()
            e.pq_wt = dev(torch.cat([w1[:, :H].t(), w1[:, H:2 * H].t()], dim=1))       # [H, 2H]
            e.pq_b = dev(torch.cat([first.bias, torch.zeros_like(first.bias)]))
            e.w_rd = dev(w1[:, 2 * H:2 * H + 2].t())                                   # [2, H]
            e.w2t = dev(second.weight.t())
            e.b2 = dev(second.bias)
            e.w_out = dev(head.weight.reshape(H)) if head is not None else dev(torch.zeros(H, device=w1.device))
            e.b_out = dev(head_bias.reshape(1)) if head_bias is not None else None
            e.tc_pack = tc_pack(second.weight)
            e.tc_pack_pq = tc_pack(torch.cat([w1[:, :H], w1[:, H:2 * H]], dim=0))      # [2H, H]: block 0 -> P, 1 -> Q
            return e

        w = _lib.EgnnWeights()
        w.emb_w, w.emb_b = dev(self.embedding.weight), dev(self.embedding.bias)
        w.out_w, w.out_b = dev(self.embedding_out.weight), dev(self.embedding_out.bias)
        for b in range(self.n_layers):
            blk = getattr(self, f"e_block_{b}")
            for s in range(self.inv_sublayers):
                g = getattr(blk, f"gcl_{s}")
                cg = w.block[b].gcl[s]
                att = g.att_mlp[0] if self.attention else None
                cg.edge = edge(g.edge_mlp[0], g.edge_mlp[2], att, att.bias if att is not None else None)
                cg.node_w1t, cg.node_b1 = dev(g.node_mlp[0].weight.t()), dev(g.node_mlp[0].bias)
                cg.node_w2t, cg.node_b2 = dev(g.node_mlp[2].weight.t()), dev(g.node_mlp[2].bias)
                cg.tc_pack_node1 = tc_pack(g.node_mlp[0].weight)
                cg.tc_pack_node2 = tc_pack(g.node_mlp[2].weight)
            q = blk.gcl_equiv
            w.block[b].equiv = edge(q.coord_mlp[0], q.coord_mlp[2], q.coord_mlp[4], None)
            w.block[b].tc_pack_pq4, w.block[b].pq4_b = None, None
            if tcore and b + 1 < self.n_layers:
                nxt = getattr(getattr(self, f"e_block_{b + 1}"), "gcl_0").edge_mlp[0]
                we, wn = q.coord_mlp[0].weight, nxt.weight
                w.block[b].tc_pack_pq4 = tc_pack(torch.cat([we[:, :H], we[:, H:2 * H], wn[:, :H], wn[:, H:2 * H]], dim=0))
                zb = torch.zeros_like(nxt.bias)
                w.block[b].pq4_b = dev(torch.cat([q.coord_mlp[0].bias, zb, nxt.bias, zb]))
        self._pack, self._pack_key = (w, keep), key
        return self._pack

    def c_config(self, n_max_padded: int = 0) -> _lib.EgnnConfig:
        if self.aggregation_method == "mean":
            if n_max_padded <= 0:
                raise ValueError("aggregation_method='mean' divides by the PADDED edge count per receiver "
                                 "(egnn_new.py:269-273); pass the padded n_max")
            agg_div = float(n_max_padded)
        else:
            agg_div = float(self.normalization_factor)
        mode = _lib.MMA_MODES[self.mma_mode] if isinstance(self.mma_mode, str) else int(self.mma_mode)
        return _lib.EgnnConfig(self.hidden_nf, self.n_layers, self.inv_sublayers, self.in_node_nf, self.out_node_nf,
                               int(bool(self.attention)), int(bool(self.tanh)), float(self.norm_constant),
                               float(self.coords_range), agg_div, mode)

    def tile_m(self) -> int:
        mode = _lib.MMA_MODES[self.mma_mode] if isinstance(self.mma_mode, str) else int(self.mma_mode)
        return TILE_M[mode]

    def workspace(self, n_node: int, device, n_edge: int = 0) -> torch.Tensor:
        cfg = self.c_config(1)
        need = _lib.lib().geoldm_egnn_workspace_bytes(C.byref(cfg), n_node, n_edge)
        if self._ws is None or self._ws.numel() < need or self._ws.device != torch.device(device):
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    @torch.no_grad()
    def forward(self, h: torch.Tensor, x: torch.Tensor, batch: RaggedBatch, h_out: Optional[torch.Tensor] = None,
                x_out: Optional[torch.Tensor] = None, dx_out: Optional[torch.Tensor] = None):
        """h [N, in_node_nf], x [N, 3] ragged-packed fp32 CUDA tensors -> (h [N, out_node_nf], x [N, 3]).
        dx_out (optional, [N, 3]) receives the accumulated coordinate displacement x_out - x at full precision."""
        if not (h.is_cuda and x.is_cuda):
            raise _lib.GeoldmError("EGNN.forward needs CUDA tensors (no CPU path)")
        N = batch.n_node
        assert h.shape == (N, self.in_node_nf) and x.shape == (N, 3), (h.shape, x.shape, N)
        h, x = h.contiguous().float(), x.contiguous().float()
        h_out = torch.empty(N, self.out_node_nf, device=h.device) if h_out is None else h_out
        x_out = torch.empty(N, 3, device=h.device) if x_out is None else x_out
        w, _keep = self.packed()
        cfg = self.c_config(batch.n_max)
        ws = self.workspace(N, h.device, batch.n_edge)
        cb = batch.c_batch(self.tile_m())
        st = torch.cuda.current_stream(h.device).cuda_stream
        _lib.check(_lib.lib().geoldm_egnn_forward(C.byref(cfg), C.byref(w), C.byref(cb), _lib.ptr(h), _lib.ptr(x),
                                                  _lib.ptr(h_out), _lib.ptr(x_out), _lib.ptr(dx_out), _lib.ptr(ws),
                                                  ws.numel(),
                                                  C.c_void_p(st)), "geoldm_egnn_forward")
        return h_out, x_out

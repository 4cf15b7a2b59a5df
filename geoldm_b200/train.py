"""Training-side (autograd) path of the EGNN wrappers — SURVEY §8 configs 2 and 5.

Companion of the fused inference kernels: the network is evaluated on the same ragged packing.  Every ``nn.Linear`` with
GEMM-sized dimensions runs on this repo's kernels in all three directions through ``torch.autograd.Function``: edge-level
forward and input-gradient GEMMs on the tcgen05 fp16-split kernel (``geoldm_linear_tc`` / ``geoldm_linear_tc_grad``, the
gradient operand pre-scaled by a power of two), node-level GEMMs on the fp32 kernel (``geoldm_linear``), weight and bias
gradients on ``geoldm_gemm_tn_bias``, accumulated straight into ``.grad`` where the parameter is registered for it
(``register_direct_grad``, training.FlatGradBuckets).  The element-wise stages of both edge MLPs (gather + first SiLU; bias +
second SiLU + attention gate + segment sum / coordinate head), ``coord2diff`` and the coordinate update run as fused
forward/backward CUDA kernels (``geoldm_train_edge_act_*``, ``geoldm_train_edge_tail_*``, ``geoldm_train_coord2diff_*``,
``geoldm_train_coord_step_*``, csrc/train.cu).  The remaining node-level element-wise ops and the 1-3-wide heads are library
ops recorded by autograd.  It shares the ``nn.Parameter`` objects of the inference modules, so optimisers, EMA and the NCCL
gradient all-reduce work unchanged.  A recompute-in-tile backward on the tensor cores is not built.

Reference: egnn/egnn_new.py:30-65,86-105,134-147,184-197; egnn/models.py:49-113,335-381.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from .packing import RaggedBatch


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# Edge-level GEMMs (M = number of edges) of the autograd path run on the tcgen05 fp16-split kernel (fp32-equivalent
# products, csrc/edge_tc16.cu DENSE mode) in the FORWARD direction, where the A operand is an O(1) activation, and in the
# input-gradient direction with the gradient operand pre-scaled by a power of two (see _LinearFn.backward); weight
# gradients stay on the fp32 FFMA kernel, as do the node-level GEMMs (a few tiles: the launch is cheaper than packing a
# weight image).  GEOLDM_TRAIN_TC=0 disables it.
_TC_MIN_ROWS = 4096


def _tc_block(n: int) -> int:
    """Column-block width H of the tensor-core kernel for n outputs (0: not usable)."""
    for hb in (256, 192, 128, 64):
        if n % hb == 0:
            return hb
    return 0


def _tc_ok(M: int, K: int, N: int) -> bool:
    import os
    return (M >= _TC_MIN_ROWS and K % 64 == 0 and _tc_block(N) > 0 and os.environ.get("GEOLDM_TRAIN_TC", "1") != "0"
            and _lib.lib().geoldm_has_tcgen05() == 1)


def _tc_square(N, K):
    return N == K and K in (64, 128, 192, 256)


def _tc_linear(x, w_nk, bias, M, K, N, want_t=False):
    """x [M,K] @ w_nk[N,K]^T (+ bias) through geoldm_tc_pack16 + geoldm_linear_tc (3xF16).  Square layers: both operand
    images (this GEMM's and, want_t, the input-gradient GEMM's) come from ONE single-block pack launch; returns (out, pack_t)."""
    L = _lib.lib()
    hb = _tc_block(N)
    st = _stream(x)
    pack = torch.empty(L.geoldm_tc_pack16_bytes(hb, N, K), dtype=torch.uint8, device=x.device)
    pack_t = None
    if _tc_square(N, K):
        if want_t:
            pack_t = torch.empty_like(pack)
        _lib.check(L.geoldm_tc_pack16_pair(K, _lib.ptr(w_nk), _lib.ptr(pack), _lib.ptr(pack_t), st), "geoldm_tc_pack16_pair")
    else:
        _lib.check(L.geoldm_tc_pack16(hb, _lib.ptr(w_nk), N, K, _lib.ptr(pack), st), "geoldm_tc_pack16(train)")
    out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    _lib.check(L.geoldm_linear_tc(hb, 16, _lib.ptr(x), K, None, 0, 1.0, _lib.ptr(pack), N // hb, _lib.ptr(bias), None, 0,
                                  _lib.ptr(out), M, st), "geoldm_linear_tc(train)")
    return out, pack_t


# Zero-filled temporaries of one optimisation step from ONE arena that is cleared with a single fill at the top of the step
# (training.FlatGradBuckets.zero() / finish() bracket the step), instead of one fill launch per temporary (~230 per step:
# scatter targets of the fused kernels, weight-gradient buffers).  The first bracketed step only measures; the arena is
# allocated after it (outside any graph capture) and serves the following steps, whose allocation sequence is the same for
# the same batch signature; anything that does not fit falls back to torch.zeros.  Tensors are handed out as independent
# aliases of the arena's storage (no view relationship, own version counters).  Everything taken from the arena is dead
# when the step ends: gradients reach the parameters through the flat buckets, never by handing one of these tensors over
# (which is why the arena is tied to FlatGradBuckets: there .grad always exists, so autograd adds instead of keeping).
class _ZeroArena:
    def __init__(self):
        self.buf, self.off, self.need, self.active, self.measuring = None, 0, 0, False, False
        self._retired = []          # outgrown buffers stay allocated: a captured graph may still write to them on replay

    @staticmethod
    def _al(n):
        return (n + 63) // 64 * 64                      # 256-byte pieces: vector accesses stay aligned

    def begin(self):
        if self.active or self.measuring:               # nested / interleaved steps: stand aside
            self.active = self.measuring = False
            return
        self.off, self.need = 0, 0
        if self.buf is not None:
            self.buf.zero_()
            self.active = True
        else:
            self.measuring = True

    def end(self, device):
        capturing = device.type == "cuda" and torch.cuda.is_current_stream_capturing()
        if self.measuring and self.need > 0 and not capturing:
            self.buf = torch.empty(self.need, dtype=torch.float32, device=device)
        elif self.active and self.need > self.buf.numel() and not capturing:
            self._retired.append(self.buf)
            self.buf = torch.empty(self.need, dtype=torch.float32, device=device)      # a larger batch came along
        self.active = self.measuring = False

    def zeros(self, n, device):
        a = self._al(n)
        self.need += a
        if self.active and self.buf.device == device and self.off + a <= self.buf.numel():
            t = torch.empty(0, dtype=torch.float32, device=device).set_(self.buf.untyped_storage(), self.off, (n,), (1,))
            self.off += a
            return t
        return torch.zeros(n, dtype=torch.float32, device=device)


_ARENAS: dict = {}


def arena_begin(device):
    if device.type == "cuda":
        _ARENAS.setdefault(device.index, _ZeroArena()).begin()


def arena_end(device):
    if device.type == "cuda" and device.index in _ARENAS:
        _ARENAS[device.index].end(device)


def _zeros(n, device):
    """n zero-filled float32 values: from the step's arena when one is open on this device, else a fresh tensor."""
    a = _ARENAS.get(device.index) if device.type == "cuda" else None
    if a is None or not (a.active or a.measuring):
        return torch.zeros(n, dtype=torch.float32, device=device)
    return a.zeros(n, device)


# Direct gradient accumulation: a parameter registered here (training.FlatGradBuckets does it for every parameter whose
# .grad is a view of a flat bucket) receives its weight / bias gradient straight from the GEMM kernel, which ACCUMULATES
# into .grad (C += A^T B), instead of a zero-filled temporary that autograd then adds to .grad: one fill and one add launch
# less per tensor.  `armed()` says whether .grad is currently valid as an accumulation target (between zero() and finish());
# `hook(param)` is what the post-accumulate hook would have done (bucket bookkeeping, overlapped all-reduce).
_DIRECT: dict = {}


def register_direct_grad(param, armed, hook):
    import weakref
    _DIRECT[id(param)] = (weakref.ref(param), armed, hook)      # weak: the registry keeps no parameter alive


def unregister_direct_grad(param):
    _DIRECT.pop(id(param), None)


def _direct_grad(t):
    """The registered entry of tensor `t` if its gradient can be accumulated in place right now, else None."""
    e = _DIRECT.get(id(t)) if t is not None else None
    if e is None or e[0]() is not t or not e[1]():
        return None
    g = t.grad
    if g is None or g.dtype != torch.float32 or not g.is_cuda or not g.is_contiguous() or g.shape != t.shape:
        return None
    return e


# max |dmpre| of the gradient tensor that _EdgeTailFn.backward just produced, offered to the input-gradient GEMM that
# consumes it (the next node of the backward graph): one slot holding (tensor, amax).  The slot keeps the tensor alive, so
# no other tensor can sit at its address while the offer stands; it is taken once or replaced by the next offer.
_AMAX_OFFER = [None]


def _offer_amax(t, amax):
    _AMAX_OFFER[0] = (t, amax)


def _take_amax(dy):
    o = _AMAX_OFFER[0]
    if o is None:
        return None
    _AMAX_OFFER[0] = None
    t, amax = o
    if t.data_ptr() == dy.data_ptr() and t.shape == dy.shape and t._version == 0 and dy.is_contiguous():
        return amax
    return None


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the geoldm_b200 GEMM kernels, x [M,K], W [N,K] (PyTorch layout): fp16-split tcgen05 for the forward
    and (square layers) the input gradient at edge-level row counts, fp32 FFMA otherwise and for dW."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        ctx.pack_t = None
        ctx.w_param = weight if id(weight) in _DIRECT else None
        ctx.b_param = bias if (bias is not None and id(bias) in _DIRECT) else None
        if x.is_cuda and _tc_ok(M, K, N):
            out, ctx.pack_t = _tc_linear(x, weight.contiguous(), bias, M, K, N, want_t=ctx.needs_input_grad[0])
            ctx.save_for_backward(x, weight)
            ctx.has_bias = bias is not None
            return out
        wt = weight.t().contiguous()                      # k-major [K][N]
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().geoldm_linear(_lib.ptr(x), K, None, 0, 1.0, _lib.ptr(wt), _lib.ptr(bias), None, 0,
                                            _lib.ptr(out), M, N, 0, _stream(x)), "geoldm_linear(fwd)")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        L = _lib.lib()
        dx = dw = db = None
        # dX = dY W.  Edge-level square layers: fp16-split tensor-core kernel with dY pre-scaled by a power of two taken
        # from its own maximum (geoldm_linear_tc_grad) - dY is a GRADIENT, 1 / batch-size small, and leaves the normal fp16
        # range unscaled (measured before the scaling existed: fine at 64 molecules, 4.7 relative error on a bias gradient
        # at 256 molecules per GPU, scripts/train_check.py).  Everything else: fp32 FFMA kernel.
        if ctx.needs_input_grad[0]:
            w = weight.contiguous()                       # [N][K] is already k-major for dX = dY W
            dx = torch.empty(M, K, device=x.device, dtype=torch.float32)
            if _tc_square(N, K) and _tc_ok(M, N, K):
                st = _stream(x)
                pack = ctx.pack_t                          # packed together with the forward operand (same weight version)
                if pack is None:
                    pack = torch.empty(L.geoldm_tc_pack16_bytes(K, K, N), dtype=torch.uint8, device=x.device)
                    _lib.check(L.geoldm_tc_pack16_pair(K, _lib.ptr(w), None, _lib.ptr(pack), st), "geoldm_tc_pack16_pair(dX)")
                amax = _take_amax(dy)                       # max |dy| already found by the kernel that produced dy?
                ready = amax is not None
                if not ready:
                    amax = torch.empty(1, dtype=torch.int32, device=x.device)
                _lib.check(L.geoldm_linear_tc_grad(K, _lib.ptr(dy), N, _lib.ptr(pack), _lib.ptr(dx), M, _lib.ptr(amax),
                                                   int(ready), st), "geoldm_linear_tc_grad(dX)")
            else:
                _lib.check(L.geoldm_linear(_lib.ptr(dy), N, None, 0, 1.0, _lib.ptr(w), None, None, 0, _lib.ptr(dx), M, K, 0,
                                           _stream(x)), "geoldm_linear(dX)")
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            # dW and the bias gradient from ONE pass over dY (geoldm_gemm_tn_bias: the k0 == 0 blocks also add up the dY
            # columns).  Targets: the parameters' own .grad when they are registered for direct accumulation (the kernel
            # adds into its output), otherwise ONE zero-filled allocation for both, returned to autograd.
            ew = _direct_grad(ctx.w_param)
            eb = _direct_grad(ctx.b_param) if want_db else None
            n_tmp = (0 if ew else N * K) + (N if (want_db and not eb) else 0)
            buf = _zeros(n_tmp, x.device) if n_tmp else None
            if ew:
                dw_t = ew[0]().grad
            else:
                dw_t = dw = buf[:N * K].view(N, K)
            db_t = None
            if want_db:
                if eb:
                    db_t = eb[0]().grad
                else:
                    db_t = db = buf[(0 if ew else N * K):]
            _lib.check(L.geoldm_gemm_tn_bias(_lib.ptr(dy), N, _lib.ptr(x), K, _lib.ptr(dw_t), K, _lib.ptr(db_t), M, N, K,
                                             _stream(x)), "geoldm_gemm_tn_bias(dW, db)")
            for e in (ew, eb):
                if e:
                    e[2](e[0]())
        elif want_db:
            db = dy.sum(0)
        return dx, dw, db


# Test seam, never set by the product: tests/ call `allow_cpu_graph_check(True)` to run THIS autograd graph (ragged
# gathers, split first layer, segment sums, loss algebra, gloo all-reduce logic) on CPU tensors with library GEMMs
# standing in for the CUDA kernels.  Without it a CPU tensor raises: there is no CPU path in the product.
_CPU_GRAPH_CHECK = False


def allow_cpu_graph_check(flag: bool = True):
    global _CPU_GRAPH_CHECK
    _CPU_GRAPH_CHECK = bool(flag)


def _require_cuda(t):
    if not t.is_cuda and not _CPU_GRAPH_CHECK:
        raise _lib.GeoldmError("geoldm_b200 has no CPU path: the training graph needs CUDA tensors")


def linear(x, weight, bias=None):
    """nn.Linear forward/backward on our kernels when the shapes are GEMM-sized; library op only for the tiny heads
    (embedding in/out, attention and coordinate heads: K or N < 16)."""
    _require_cuda(x)
    K, N = weight.shape[1], weight.shape[0]
    if x.is_cuda and K % 16 == 0 and N % 16 == 0 and x.dtype == torch.float32:
        return _LinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class _EdgeActFn(torch.autograd.Function):
    """a[e] = SiLU(P[i_e] + Q[j_e] + r_e w_r + d0_e w_d) on `geoldm_train_edge_act_*` (forward and backward fused,
    the pre-activation is recomputed in the backward pass)."""

    @staticmethod
    def forward(ctx, pq, r, d0, w_rd, ei32, ej32):
        pq, r, d0, w_rd = pq.contiguous(), r.contiguous(), d0.contiguous(), w_rd.contiguous()
        E, H = ei32.numel(), w_rd.shape[1]
        a = torch.empty(E, H, device=pq.device, dtype=torch.float32)
        _lib.check(_lib.lib().geoldm_train_edge_act_fwd(E, H, _lib.ptr(pq), pq.shape[1], _lib.ptr(r), _lib.ptr(d0),
                                                        _lib.ptr(w_rd), _lib.ptr(ei32), _lib.ptr(ej32), _lib.ptr(a),
                                                        _stream(pq)), "train_edge_act_fwd")
        ctx.save_for_backward(pq, r, d0, w_rd, ei32, ej32)
        return a

    @staticmethod
    def backward(ctx, da):
        pq, r, d0, w_rd, ei32, ej32 = ctx.saved_tensors
        E, H = ei32.numel(), w_rd.shape[1]
        da = da.contiguous()
        buf = _zeros(pq.numel() + w_rd.numel(), pq.device)                                    # one fill for both
        dpq = buf[:pq.numel()].view_as(pq)
        dw = buf[pq.numel():].view_as(w_rd)
        dr = torch.empty(E, device=pq.device, dtype=torch.float32)
        dd0 = torch.empty(E, device=pq.device, dtype=torch.float32)
        _lib.check(_lib.lib().geoldm_train_edge_act_bwd(E, H, _lib.ptr(pq), pq.shape[1], _lib.ptr(r), _lib.ptr(d0),
                                                        _lib.ptr(w_rd), _lib.ptr(ei32), _lib.ptr(ej32), _lib.ptr(da),
                                                        _lib.ptr(dpq), _lib.ptr(dr), _lib.ptr(dd0), _lib.ptr(dw),
                                                        _stream(pq)), "train_edge_act_bwd")
        return dpq, dr.view_as(r), dd0.view_as(d0), dw, None, None


class _EdgeTailFn(torch.autograd.Function):
    """m = SiLU(mpre + b2) followed by the gated segment sum (gate=True -> agg [N, H]) or the head dot product
    (gate=False -> sc [E, 1]) on `geoldm_train_edge_tail_*`."""

    @staticmethod
    def forward(ctx, mpre, b2, w, bw, ei32, n_node, div, gate, attention):
        mpre, b2 = mpre.contiguous(), b2.contiguous()
        E, H = mpre.shape
        w = None if w is None else w.contiguous()
        dev = mpre.device
        agg = _zeros(n_node * H, dev).view(n_node, H) if gate else None
        sc = None if gate else torch.empty(E, 1, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().geoldm_train_edge_tail_fwd(E, H, _lib.ptr(mpre), _lib.ptr(b2), _lib.ptr(w), _lib.ptr(bw),
                                                         int(gate), int(attention), _lib.ptr(ei32), float(div),
                                                         _lib.ptr(agg), _lib.ptr(sc), _stream(mpre)), "train_edge_tail_fwd")
        ctx.save_for_backward(mpre, b2, w, bw, ei32)
        ctx.cfg = (float(div), bool(gate), bool(attention))
        return agg if gate else sc

    @staticmethod
    def backward(ctx, dout):
        mpre, b2, w, bw, ei32 = ctx.saved_tensors
        div, gate, attention = ctx.cfg
        E, H = mpre.shape
        dout = dout.contiguous()
        dmpre = torch.empty_like(mpre)
        # db2 | dw | dbw carved from one zero-filled allocation (H is a multiple of 4: the vector pieces stay 16-byte aligned)
        buf = _zeros(2 * H + 4, mpre.device)
        db2 = buf[:H]
        dw = None if w is None else buf[H:2 * H]
        dbw = None if bw is None else buf[2 * H:2 * H + 1]
        amax = buf[2 * H + 1:2 * H + 2]                  # bit pattern of max |dmpre| (0.0f = 0u to start with)
        # scratch of the deterministic double-precision reduction of the attention-bias gradient: per-block partials + a
        # counter the kernel leaves at zero, so ONE zero-initialised buffer per (device, grid) serves every launch
        scratch = None
        if dbw is not None and gate and attention:
            scratch = _bw_scratch(mpre.device, _lib.lib().geoldm_train_bwd_blocks(E))
        _lib.check(_lib.lib().geoldm_train_edge_tail_bwd(E, H, _lib.ptr(mpre), _lib.ptr(b2), _lib.ptr(w), _lib.ptr(bw),
                                                         int(gate), int(attention), _lib.ptr(ei32), div,
                                                         _lib.ptr(dout) if gate else None, None if gate else _lib.ptr(dout),
                                                         _lib.ptr(dmpre), _lib.ptr(db2), _lib.ptr(dw), _lib.ptr(dbw),
                                                         _lib.ptr(scratch), _lib.ptr(amax), _stream(mpre)), "train_edge_tail_bwd")
        _offer_amax(dmpre, amax)
        if w is not None and not (attention or not gate):
            dw = None
        return dmpre, db2, dw, dbw, None, None, None, None, None


_BW_SCRATCH: dict = {}


def _bw_scratch(dev, n_blocks):
    """[n_blocks + 1] float64 partials + arrival counter of geoldm_train_edge_tail_bwd (launches on one stream run in order,
    every launch overwrites all partials and resets the counter)."""
    key = (dev.index, int(n_blocks))
    t = _BW_SCRATCH.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.GeoldmError("the attention-bias scratch must exist before graph capture: run one eager step first")
        t = _BW_SCRATCH[key] = torch.zeros(n_blocks + 1, dtype=torch.float64, device=dev)
    return t


class _SplitFirstFn(torch.autograd.Function):
    """First edge layer weight [H, 2H + 2] / bias [H] -> the operands of its split form: projection weight [2H, H] (P rows,
    then Q rows), projection bias [2H] (the layer bias on the P half, zero on the Q half) and the two distance columns as
    [2, H].  One concatenation forward and one backward instead of autograd's slice / pad / accumulate chain per piece."""

    @staticmethod
    def forward(ctx, w1, b1):
        H = w1.shape[0]
        wpq = torch.cat([w1[:, :H], w1[:, H:2 * H]], dim=0)
        bpq = torch.nn.functional.pad(b1, (0, H))
        w_rd = w1[:, 2 * H:2 * H + 2].t().contiguous()
        ctx.shape = tuple(w1.shape)
        return wpq, bpq, w_rd

    @staticmethod
    def backward(ctx, dwpq, dbpq, dw_rd):
        H, C = ctx.shape
        ref = dwpq if dwpq is not None else dw_rd if dw_rd is not None else dbpq
        z = lambda *s: torch.zeros(*s, device=ref.device, dtype=ref.dtype)
        dwpq = z(2 * H, H) if dwpq is None else dwpq
        dw_rd = z(2, H) if dw_rd is None else dw_rd
        parts = [dwpq[:H], dwpq[H:], dw_rd.t()]
        if C > 2 * H + 2:
            parts.append(z(H, C - 2 * H - 2))
        return torch.cat(parts, dim=1), (None if dbpq is None else dbpq[:H])


class _Coord2DiffFn(torch.autograd.Function):
    """coord2diff (egnn_new.py:249-255) as one launch forward and one backward: x [N, 3] -> r [E, 1] = |x_i - x_j|^2 and
    (want_u) u [E, 3] = (x_i - x_j) / (sqrt(r + 1e-8) + norm_constant)."""

    @staticmethod
    def forward(ctx, x, ei32, ej32, norm_constant, want_u):
        xc = x.contiguous()
        E = ei32.numel()
        r = torch.empty(E, 1, device=x.device, dtype=torch.float32)
        u = torch.empty(E, 3, device=x.device, dtype=torch.float32) if want_u else None
        _lib.check(_lib.lib().geoldm_train_coord2diff_fwd(E, _lib.ptr(xc), _lib.ptr(ei32), _lib.ptr(ej32),
                                                          float(norm_constant), _lib.ptr(r), _lib.ptr(u), _stream(x)),
                   "train_coord2diff_fwd")
        ctx.save_for_backward(xc, ei32, ej32)
        ctx.c = float(norm_constant)
        ctx.set_materialize_grads(False)
        return r, u

    @staticmethod
    def backward(ctx, gr, gu):
        xc, ei32, ej32 = ctx.saved_tensors
        gx = _zeros(xc.numel(), xc.device).view_as(xc)
        gr = None if gr is None else gr.contiguous()
        gu = None if gu is None else gu.contiguous()
        _lib.check(_lib.lib().geoldm_train_coord2diff_bwd(ei32.numel(), _lib.ptr(xc), _lib.ptr(ei32), _lib.ptr(ej32), ctx.c,
                                                          _lib.ptr(gr), _lib.ptr(gu), _lib.ptr(gx), _stream(xc)),
                   "train_coord2diff_bwd")
        return gx, None, None, None, None


class _CoordStepFn(torch.autograd.Function):
    """Coordinate update of EquivariantUpdate (egnn_new.py:91-99): step [N, 3] = segment sum over receivers of
    u * (tanh(sc) * coords_range | sc) / normalization_factor, one launch forward and one backward."""

    @staticmethod
    def forward(ctx, u, sc, ei32, n_node, use_tanh, coords_range, div):
        u, sc = u.contiguous(), sc.contiguous()
        step = _zeros(n_node * 3, u.device).view(n_node, 3)
        ctx.cfg = (int(bool(use_tanh)), float(coords_range), float(div))
        _lib.check(_lib.lib().geoldm_train_coord_step_fwd(ei32.numel(), _lib.ptr(u), _lib.ptr(sc), _lib.ptr(ei32), *ctx.cfg,
                                                          _lib.ptr(step), _stream(u)), "train_coord_step_fwd")
        ctx.save_for_backward(u, sc, ei32)
        return step

    @staticmethod
    def backward(ctx, gstep):
        u, sc, ei32 = ctx.saved_tensors
        gstep = gstep.contiguous()
        gu, gsc = torch.empty_like(u), torch.empty_like(sc)
        _lib.check(_lib.lib().geoldm_train_coord_step_bwd(ei32.numel(), _lib.ptr(u), _lib.ptr(sc), _lib.ptr(ei32), *ctx.cfg,
                                                          _lib.ptr(gstep), _lib.ptr(gu), _lib.ptr(gsc), _stream(u)),
                   "train_coord_step_bwd")
        return gu, gsc, None, None, None, None, None


def _fused_ok(h, H):
    return h.is_cuda and h.dtype == torch.float32 and H <= 256 and H % 16 == 0


def _coord2diff(x, ei, ej, norm_constant):
    d = x.index_select(0, ei) - x.index_select(0, ej)
    r = (d * d).sum(1, keepdim=True)
    return r, d / (torch.sqrt(r + 1e-8) + norm_constant)


def _edge_pre(h, first, ei, ej, r, d0, H, e32):
    """First edge layer in split form: per-node projections + the two distance columns, then SiLU."""
    w1 = first.weight
    wpq, bpq, w_rd = _SplitFirstFn.apply(w1, first.bias)                    # [2H, H], [2H], [2, H]
    pq = linear(h, wpq, bpq)
    if _fused_ok(h, H):
        return _EdgeActFn.apply(pq, r.reshape(-1), d0.reshape(-1), w_rd, e32[0], e32[1])
    pre1 = pq[:, :H].index_select(0, ei) + pq[:, H:].index_select(0, ej) + r * w1[:, 2 * H] + d0 * w1[:, 2 * H + 1]
    return F.silu(pre1)


def egnn_forward_train(egnn, h, x, batch: RaggedBatch):
    """EGNN.forward with autograd on ragged tensors: h [N, in_nf], x [N, 3] -> (h [N, out_nf], x [N, 3])."""
    H = egnn.hidden_nf
    ei, ej = batch.edge_i.long(), batch.edge_j.long()
    e32 = (batch.edge_i.to(torch.int32).contiguous(), batch.edge_j.to(torch.int32).contiguous())
    N = batch.n_node
    if egnn.aggregation_method == "mean":
        if batch.n_max <= 0:
            raise ValueError("aggregation_method='mean' needs the padded n_max (egnn_new.py:269-273)")
        div = float(batch.n_max)
    else:
        div = float(egnn.normalization_factor)
    fused = _fused_ok(h, H) and x.is_cuda and x.dtype == torch.float32
    if fused:
        d0, _ = _Coord2DiffFn.apply(x, e32[0], e32[1], 1.0, False)
    else:
        d0, _ = _coord2diff(x, ei, ej, 1.0)
    x0, dx = x, None        # the accumulated displacement is carried next to x: the dynamics' velocity x_final - x is
    h = F.linear(h, egnn.embedding.weight, egnn.embedding.bias)   # then exact instead of a cancelling difference
    for b in range(egnn.n_layers):
        blk = getattr(egnn, f"e_block_{b}")
        if fused:
            r, u = _Coord2DiffFn.apply(x, e32[0], e32[1], float(egnn.norm_constant), True)
        else:
            r, u = _coord2diff(x, ei, ej, float(egnn.norm_constant))
        for s in range(egnn.inv_sublayers):
            g = getattr(blk, f"gcl_{s}")
            a1 = _edge_pre(h, g.edge_mlp[0], ei, ej, r, d0, H, e32)
            if fused:      # second layer without bias; bias + SiLU + gate + segment sum in one kernel
                att_w = g.att_mlp[0].weight.reshape(H) if egnn.attention else None
                att_b = g.att_mlp[0].bias if egnn.attention else None
                agg = _EdgeTailFn.apply(linear(a1, g.edge_mlp[2].weight, None), g.edge_mlp[2].bias, att_w, att_b, e32[0],
                                        N, div, True, bool(egnn.attention))
            else:
                m = F.silu(linear(a1, g.edge_mlp[2].weight, g.edge_mlp[2].bias))
                if egnn.attention:
                    m = m * torch.sigmoid(F.linear(m, g.att_mlp[0].weight, g.att_mlp[0].bias))
                agg = torch.zeros(N, H, device=h.device, dtype=h.dtype).index_add_(0, ei, m) / div
            t1 = F.silu(linear(torch.cat([h, agg], dim=1), g.node_mlp[0].weight, g.node_mlp[0].bias))
            h = h + linear(t1, g.node_mlp[2].weight, g.node_mlp[2].bias)
        q = blk.gcl_equiv
        a2 = _edge_pre(h, q.coord_mlp[0], ei, ej, r, d0, H, e32)
        if fused:
            sc = _EdgeTailFn.apply(linear(a2, q.coord_mlp[2].weight, None), q.coord_mlp[2].bias,
                                   q.coord_mlp[4].weight.reshape(H), None, e32[0], N, div, False, False)
        else:
            m2 = F.silu(linear(a2, q.coord_mlp[2].weight, q.coord_mlp[2].bias))
            sc = F.linear(m2, q.coord_mlp[4].weight)
        if fused:
            step = _CoordStepFn.apply(u, sc, e32[0], N, bool(egnn.tanh), float(egnn.coords_range), div)
        else:
            trans = u * torch.tanh(sc) * egnn.coords_range if egnn.tanh else u * sc
            step = torch.zeros(N, 3, device=x.device, dtype=x.dtype).index_add_(0, ei, trans) / div
        dx = step if dx is None else dx + step
        x = x0 + dx
    h = F.linear(h, egnn.embedding_out.weight, egnn.embedding_out.bias)
    return h, x, dx


def _remove_mean_ragged(v, batch: RaggedBatch):
    mol = batch.node_mol.long()
    cnt = (batch.mol_off[1:] - batch.mol_off[:-1]).to(v.dtype).unsqueeze(1)          # device table: no host copy
    mean = torch.zeros(batch.n_mol, v.shape[1], device=v.device, dtype=v.dtype).index_add_(0, mol, v) / cnt
    return v - mean.index_select(0, mol)


def wrapper_forward_train(wrapper, t, xh, node_mask, edge_mask, context, is_dynamics: bool):
    """EGNN_dynamics_QM9._forward / EGNN_decoder_QM9._forward with autograd (padded in, padded out)."""
    _require_cuda(xh)
    bs, n_nodes, dims = xh.shape
    batch = wrapper._masks.get(node_mask.view(bs, n_nodes, 1), edge_mask, wrapper.validate_masks)
    src = batch.node_src.long()
    flat = xh.reshape(bs * n_nodes, dims).index_select(0, src)
    x = flat[:, :wrapper.n_dims]
    h = flat[:, wrapper.n_dims:]
    if is_dynamics and wrapper.condition_time:
        t = torch.as_tensor(t, dtype=xh.dtype, device=xh.device)
        t_mol = t.reshape(1).expand(bs) if t.numel() == 1 else t.reshape(bs)
        h = torch.cat([h, t_mol.index_select(0, batch.node_mol.long()).unsqueeze(1)], dim=1)
    if context is not None:
        h = torch.cat([h, context.reshape(bs * n_nodes, -1).index_select(0, src)], dim=1)
    h_f, x_f, dx = egnn_forward_train(wrapper.egnn, h, x, batch)
    if is_dynamics:
        vel = dx if dx is not None else x_f - x
        keep = wrapper.egnn.out_node_nf - wrapper.context_node_nf - int(wrapper.condition_time)
        h_f = h_f[:, :keep]
    else:
        vel = x_f
    vel = torch.where(torch.isnan(vel).any(), torch.zeros_like(vel), vel)      # batch-wide NaN guard (models.py:100-102)
    vel = _remove_mean_ragged(vel, batch)
    out_r = torch.cat([vel, h_f], dim=1)
    out = torch.zeros(bs * n_nodes, out_r.shape[1], device=xh.device, dtype=xh.dtype).index_copy(0, src, out_r)
    out = out.view(bs, n_nodes, -1)
    if is_dynamics:
        return out
    return out[:, :, :wrapper.n_dims], out[:, :, wrapper.n_dims:]

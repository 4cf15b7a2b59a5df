"""One optimisation step of GeoLDM on one or several GPUs (BASELINE config 5).

Host-side mirror of the reference's training inner loop — train_test.py:15-70 (zero_grad, compute_loss_and_nll, backward,
adaptive clipping, AdamW step, EMA), utils.py:5-66 (EMA, Queue, gradient_clipping) and qm9/models.py:169-175 (optimiser)
— with the reference's single-process ``DataParallel`` replaced by one process per GPU: every rank evaluates its own
shard of the batch and gradients are averaged with NCCL all-reduces over flat buckets (dynamics / decoder) before the
global-norm clipping, so all ranks apply the identical update.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import losses, train


class EMA:
    """Exponential moving average of the weights (utils.py:5-28)."""

    def __init__(self, beta):
        self.beta = beta

    def update_average(self, old, new):
        return new if old is None else old * self.beta + (1 - self.beta) * new

    @torch.no_grad()
    def update_model_average(self, ma_model, current_model):
        cur = list(current_model.parameters())
        avg = list(ma_model.parameters())
        torch._foreach_mul_(avg, self.beta)
        torch._foreach_add_(avg, cur, alpha=1 - self.beta)


class Queue:
    """Last max_len gradient norms (utils.py:29-47)."""

    def __init__(self, max_len=50):
        self.items: List[float] = []
        self.max_len = max_len

    def __len__(self):
        return len(self.items)

    def add(self, item):
        self.items.insert(0, item)
        if len(self.items) > self.max_len:
            self.items.pop()

    def mean(self):
        return np.mean(self.items)

    def std(self):
        return np.std(self.items)


def gradient_clipping(flow, gradnorm_queue: Queue):
    """Clip the global gradient norm to 1.5 * mean + 2 * std of the recent history (utils.py:50-66)."""
    max_grad_norm = 1.5 * gradnorm_queue.mean() + 2 * gradnorm_queue.std()
    grad_norm = torch.nn.utils.clip_grad_norm_(flow.parameters(), max_norm=max_grad_norm, norm_type=2.0)
    gn = float(grad_norm)
    gradnorm_queue.add(float(max_grad_norm) if gn > max_grad_norm else gn)
    return grad_norm


def get_optim(args, generative_model, capturable: bool = False):
    """qm9/models.py:169-175.  `capturable=True`: step counters live on the device, so that the optimiser step can be part
    of a captured CUDA graph (GraphedTrainStep)."""
    if capturable:
        # fused multi-tensor implementation (same update rule; one kernel per parameter chunk instead of a chain of
        # element-wise launches for the bias corrections): 2.7 -> 0.4 ms of the captured step
        return torch.optim.AdamW(generative_model.parameters(), lr=args.lr, amsgrad=True, weight_decay=1e-12,
                                 capturable=True, fused=True)
    return torch.optim.AdamW(generative_model.parameters(), lr=args.lr, amsgrad=True, weight_decay=1e-12)


class FusedAdamWEMA:
    """`optim.step()` + `EMA.update_model_average` of a captured step as ONE launch of geoldm_adamw_ema_step over a device
    table of (parameter, gradient, exp_avg, exp_avg_sq, max_exp_avg_sq, EMA copy, step) pointers.  The torch optimiser stays
    the owner of the state: the kernel updates ITS tensors in place, so `optim.state_dict()` (the reference's checkpoint
    format) keeps working, and a later eager `optim.step()` continues from the same moments.  Needs an AdamW built by
    `get_optim(..., capturable=True)` whose state exists (one eager `optim.step()` has run) and float32 contiguous tensors.
    Hyper-parameters are read when `step()` is called (during capture: baked into the graph, like the library's)."""

    def __init__(self, optim, model, model_ema=None, ema: Optional[EMA] = None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        groups = optim.param_groups
        if len(groups) != 1:
            raise ValueError("FusedAdamWEMA: one parameter group expected (qm9/models.py:169-175)")
        self.group = groups[0]
        params = [p for p in self.group["params"] if p.requires_grad]
        ema_of = {}
        if model_ema is not None and ema is not None:
            ema_of = {id(p): q for p, q in zip(model.parameters(), model_ema.parameters())}
        self.ema_beta = float(ema.beta) if (ema is not None and model_ema is not None) else 1.0
        self.amsgrad = bool(self.group.get("amsgrad", False))
        dev = params[0].device
        chunk = int(_lib.lib().geoldm_optim_chunk())
        rows, cmap, self._keep = [], [], []
        for p in params:
            st = optim.state.get(p)
            if not st or "exp_avg" not in st or not torch.is_tensor(st.get("step")) or not st["step"].is_cuda:
                raise ValueError("FusedAdamWEMA: optimiser state missing (run one eager step of a capturable AdamW first)")
            if p.grad is None:
                raise ValueError("FusedAdamWEMA: every parameter needs a .grad tensor (FlatGradBuckets)")
            tensors = [p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], st.get("max_exp_avg_sq") if self.amsgrad else None,
                       ema_of.get(id(p)), st["step"]]
            for t in tensors:
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev):
                    raise ValueError("FusedAdamWEMA: float32 contiguous tensors on one device expected")
            self._keep.append(tensors)
            ptrs = [0 if t is None else t.data_ptr() for t in tensors]
            ti = len(rows)
            rows.append(ptrs + [p.numel()])
            cmap.extend((ti, off) for off in range(0, p.numel(), chunk))
        # device table: 7 pointers + (int n, int pad) = 8 x 8 bytes per tensor
        tab = torch.tensor([r[:7] + [r[7]] for r in rows], dtype=torch.int64)       # n in the low half of the last word
        self.table = tab.to(dev)
        self.chunk_map = torch.tensor(cmap, dtype=torch.int32).to(dev)
        self.n_chunks = len(cmap)
        self.step_count = self._keep[0][6].detach().clone().reshape(())             # shared update counter (device scalar)
        self.grad_ptrs = [r[1] for r in rows]

    def check_attached(self, optim):
        """The gradient views must still be the tensors whose addresses are in the table."""
        for tensors, ptr in zip(self._keep, self.grad_ptrs):
            if tensors[1].data_ptr() != ptr:
                raise RuntimeError("FusedAdamWEMA: a gradient tensor moved")

    @torch.no_grad()
    def step(self, grad_scale=None):
        """One update.  grad_scale: optional device scalar multiplied into the gradients on the fly."""
        g = self.group
        self.step_count.add_(1.0)
        stream = self._C.c_void_p(torch.cuda.current_stream(self.table.device).cuda_stream)
        b1, b2 = g["betas"]
        self._lib.check(self._lib.lib().geoldm_adamw_ema_step(
            self._lib.ptr(self.table), self._lib.ptr(self.chunk_map), self.n_chunks, self._lib.ptr(self.step_count),
            self._lib.ptr(grad_scale), float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
            int(self.amsgrad), float(self.ema_beta), stream), "geoldm_adamw_ema_step")


class DeviceGradClip:
    """utils.py:29-66 (Queue + gradient_clipping) with the history on the DEVICE: the last 50 gradient norms in a ring
    buffer, threshold 1.5 mean + 2 std (population std, as np.std), clip coefficient max_norm / (norm + 1e-6) clamped to 1
    (torch.nn.utils.clip_grad_norm_), and the queue receives the threshold instead of the norm when clipping happened.
    No host read: usable inside a captured CUDA graph."""

    def __init__(self, device, max_len=50, first=3000.0):
        self.hist = torch.zeros(max_len, device=device)
        self.count = torch.zeros((), dtype=torch.long, device=device)
        self.pos = torch.zeros((), dtype=torch.long, device=device)
        self.max_len = max_len
        self._idx = torch.arange(max_len, device=device)
        self.add(torch.tensor(float(first), device=device))

    def add(self, value):
        self.hist.index_copy_(0, self.pos.reshape(1), value.reshape(1).to(self.hist.dtype))
        self.pos.copy_((self.pos + 1) % self.max_len)
        self.count.copy_(torch.clamp(self.count + 1, max=self.max_len))

    def threshold(self):
        valid = (self._idx < self.count).to(self.hist.dtype)
        n = self.count.to(self.hist.dtype)
        mean = (self.hist * valid).sum() / n
        var = (((self.hist - mean) ** 2) * valid).sum() / n
        return 1.5 * mean + 2.0 * var.sqrt()

    @torch.no_grad()
    def clip_(self, grads):
        """grads: list of gradient tensors (e.g. the flat bucket buffers).  Returns the total norm (device scalar)."""
        max_norm = self.threshold()
        norms = torch._foreach_norm(grads, 2.0)
        total = torch.linalg.vector_norm(torch.stack(norms), 2.0)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        torch._foreach_mul_(grads, coef)
        self.add(torch.where(total > max_norm, max_norm, total))
        return total


def gradient_buckets(model) -> List[List[torch.nn.Parameter]]:
    """EVERY parameter that can receive a gradient, grouped per sub-network, sub-networks and their parameters in
    reverse registration order (the order in which backward produces them): for the latent model the decoder first,
    then the denoiser (its encoder is frozen inside `forward`: en_diffusion.py:1155, and carries no gradient unless
    the first stage is trainable); for a first-stage EnHierarchicalVAE the decoder, then the encoder.  A parameter with
    requires_grad that is in no bucket would make the ranks diverge silently, so every one lands in a bucket."""
    groups: dict = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        parts = name.split(".")
        key = ".".join(parts[:2]) if parts[0] == "vae" and len(parts) > 2 else parts[0]
        groups.setdefault(key, []).append(p)
    return [list(reversed(g)) for g in reversed(list(groups.values())) if g]


def allreduce_gradients(buckets: Iterable[List[torch.nn.Parameter]], group=None) -> int:
    """Average gradients over ranks: one flat all-reduce per bucket.  Returns the number of bytes reduced.
    Parameters whose grad is None on this rank contribute zeros (every rank must issue the same collectives)."""
    world = dist.get_world_size(group)
    total = 0
    for params in buckets:
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        for p, g in zip(params, torch._utils._unflatten_dense_tensors(flat, grads)):
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
        total += flat.numel() * flat.element_size()
    return total


class FlatGradBuckets:
    """Gradient buckets whose ``.grad`` tensors are VIEWS of one flat buffer per bucket (buckets in reverse
    registration order, overlapped with backward).  Zeroing is one memset per bucket, the all-reduce runs on the flat
    buffer itself (no flatten / unflatten copies), and a bucket's all-reduce is launched asynchronously from a
    post-accumulate hook as soon as its last gradient has been produced, i.e. while backward is still working on the
    earlier layers; ``finish`` waits for the collectives and averages.  The same object serves a single process
    (no collective, only the cheap zeroing)."""

    def __init__(self, model, group=None):
        self.group = group
        self.buckets = gradient_buckets(model)
        self.flat, self._pending, self._works, self._hooks, self._direct = [], [], [], [], []
        self._armed = False
        import weakref
        me = weakref.ref(self)
        armed = lambda: bool(me() is not None and me()._armed)
        for bi, params in enumerate(self.buckets):
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
            off = 0
            for p in params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                hook = self._make_hook(bi)
                self._hooks.append(p.register_post_accumulate_grad_hook(hook))
                # the GEMM kernels may add a weight / bias gradient straight into the view (train._LinearFn.backward)
                train.register_direct_grad(p, armed, lambda param, bi=bi: me() is not None and me()._make_hook(bi)(param))
                self._direct.append(p)
            self.flat.append(flat)
            self._pending.append(0)
        self.nbytes = sum(f.numel() * f.element_size() for f in self.flat)
        self._armed = False

    def __iter__(self):          # usable wherever a list of parameter lists is expected
        return iter(self.buckets)

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, bi):
        def hook(param):
            if not self._armed:
                return
            self._pending[bi] -= 1
            if self._pending[bi] == 0 and self._distributed():
                self._works.append(dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return hook

    def zero(self):
        """Replaces optim.zero_grad(): the views stay attached, the flat buffers are cleared."""
        for bi, (flat, params) in enumerate(zip(self.flat, self.buckets)):
            flat.zero_()
            off = 0
            for p in params:                  # re-attach views an optimiser / user may have dropped (set_to_none)
                if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + off * flat.element_size():
                    p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self._pending[bi] = len(params)
        self._works = []
        self._armed = True
        if self.flat:
            train.arena_begin(self.flat[0].device)    # zero-filled temporaries of this step: one arena, one fill

    def finish(self) -> int:
        """After backward: launch what the hooks could not (parameters that received no gradient on this rank keep
        their zeros; every rank issues the same collectives), wait, average.  Returns the bytes all-reduced."""
        self._armed = False
        if self.flat:
            train.arena_end(self.flat[0].device)
        if not self._distributed():
            return 0
        for bi, flat in enumerate(self.flat):
            if self._pending[bi] != 0:
                self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                self._pending[bi] = 0
        for w in self._works:
            w.wait()
        world = dist.get_world_size(self.group)
        for flat in self.flat:
            flat.div_(world)
        self._works = []
        return self.nbytes

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p in self._direct:
            train.unregister_direct_grad(p)
        self._direct = []


def train_step(args, model, optim, nodes_dist, x, h, node_mask, edge_mask, context, *, gradnorm_queue: Optional[Queue],
               model_ema=None, ema: Optional[EMA] = None, buckets=None, group=None, draws=None):
    """One iteration of train_test.py:train_epoch on this rank's shard.  Returns (nll, grad_norm)."""
    model.train()
    flat = buckets if isinstance(buckets, FlatGradBuckets) else None
    if flat is not None:
        flat.zero()                       # gradients live in flat per-bucket buffers; all-reduce overlaps backward
    else:
        optim.zero_grad(set_to_none=True)
    nll, reg_term, _ = losses.compute_loss_and_nll(args, model, nodes_dist, x, h, node_mask, edge_mask, context,
                                                   draws=draws)
    loss = nll + getattr(args, "ode_regularization", 0.0) * reg_term.squeeze()
    loss.backward()
    if flat is not None:
        flat.finish()
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        allreduce_gradients(buckets if buckets is not None else gradient_buckets(model), group)
    grad_norm = 0.
    if getattr(args, "clip_grad", True) and gradnorm_queue is not None:
        grad_norm = gradient_clipping(model, gradnorm_queue)
    optim.step()
    if ema is not None and model_ema is not None and getattr(args, "ema_decay", 0) > 0:
        ema.update_model_average(model_ema, model)
    return nll.detach(), grad_norm


class GraphedTrainStep:
    """train_step (zero, loss, gradients, bucketed all-reduce, adaptive clipping, AdamW, EMA) captured ONCE as a CUDA graph
    for a fixed batch signature (the molecule sizes, i.e. node_mask / edge_mask) and replayed per step: the ~3000 kernel
    launches of the eager step cost one graph launch, no host synchronisation happens inside the step (the gradient-norm
    history lives on the device: DeviceGradClip), and the step's inputs are copied into static buffers.  Batches with
    another signature need their own instance (or the eager `train_step`).

    Autograd binds a parameter's AccumulateGrad node to the stream on which it was created; the warm-up here runs on the
    side stream the capture uses, so fresh nodes are fine.  An autograd graph from an EARLIER backward on the default stream
    that is still referenced (e.g. a loss tensor that was not detached) keeps legacy-stream nodes alive and makes the capture
    fail with cudaErrorStreamCaptureImplicit: drop such references first.

    The optimiser must have been built with `capturable=True` (get_optim(..., capturable=True))."""

    def __init__(self, args, model, optim, nodes_dist, x, h, node_mask, edge_mask, context, *, model_ema=None,
                 ema: Optional[EMA] = None, buckets: Optional[FlatGradBuckets] = None, clip: Optional[DeviceGradClip] = None,
                 warmup: int = 3, fused_optim: bool = True):
        dev = x.device
        self.fused_optim = None
        self.args, self.model, self.optim, self.nodes_dist = args, model, optim, nodes_dist
        self.model_ema, self.ema = model_ema, ema
        self.buckets = buckets if buckets is not None else FlatGradBuckets(model)
        self.clip = clip if clip is not None else (DeviceGradClip(dev) if getattr(args, "clip_grad", True) else None)
        self.x = x.clone()
        self.h = {k: v.clone() for k, v in h.items()}
        self.context = None if context is None else context.clone()
        self.node_mask, self.edge_mask = node_mask, edge_mask
        if float((x * (1 - node_mask)).abs().sum()) >= 1e-8:          # the step itself must not read back (losses.py)
            raise AssertionError("x is not masked")
        model.train()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        with torch.cuda.stream(side):                                  # eager warm-up on the stream the capture will use
            for _ in range(max(1, warmup)):                            # >= 1: mask packing and lazy state need host work
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # optimiser + EMA of the captured step as one multi-tensor launch (the warm-up above ran the library optimiser,
        # which created its state); GEOLDM_FUSED_OPTIM=0 keeps the library's fused optimiser and the foreach EMA
        self.fused_optim = None
        import os
        if fused_optim and os.environ.get("GEOLDM_FUSED_OPTIM", "1") != "0":
            use_ema = self.ema is not None and self.model_ema is not None and getattr(self.args, "ema_decay", 0) > 0
            self.fused_optim = FusedAdamWEMA(optim, model, self.model_ema if use_ema else None, self.ema if use_ema else None)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.nll, self.grad_norm = self._body()

    def _body(self):
        self.buckets.zero()
        nll, reg_term, _ = losses.compute_loss_and_nll(self.args, self.model, self.nodes_dist, self.x, self.h,
                                                       self.node_mask, self.edge_mask, self.context)
        loss = nll + getattr(self.args, "ode_regularization", 0.0) * reg_term.squeeze()
        loss.backward()
        self.buckets.finish()
        grad_norm = self.clip.clip_(self.buckets.flat) if self.clip is not None else torch.zeros((), device=nll.device)
        if self.fused_optim is not None:
            self.fused_optim.check_attached(self.optim)
            self.fused_optim.step()
        else:
            self.optim.step()
            if self.ema is not None and self.model_ema is not None and getattr(self.args, "ema_decay", 0) > 0:
                self.ema.update_model_average(self.model_ema, self.model)
        return nll.detach(), grad_norm

    def close(self):
        """Release the captured graph (it holds NCCL work when the step all-reduces: destroy it BEFORE
        torch.distributed.destroy_process_group, which otherwise waits forever)."""
        torch.cuda.synchronize()
        self.graph = None
        self.nll = self.grad_norm = None

    def __call__(self, x, h=None, context=None):
        """Copies the step's inputs into the static buffers (asynchronously) and replays the graph.  Returns device
        scalars (nll, gradient norm) that the next call overwrites."""
        self.x.copy_(x, non_blocking=True)
        if h is not None:
            for k, v in h.items():
                if v.numel():
                    self.h[k].copy_(v, non_blocking=True)
        if context is not None and self.context is not None:
            self.context.copy_(context, non_blocking=True)
        self.graph.replay()
        for net in (self.model, self.model_ema):       # weights changed without python noticing: drop packed weight images
            if net is None:
                continue
            for m in net.modules():
                if hasattr(m, "_pack_key"):
                    m._pack_key = None
        return self.nll, self.grad_norm

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

from . import losses


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


def get_optim(args, generative_model):
    """qm9/models.py:169-175."""
    return torch.optim.AdamW(generative_model.parameters(), lr=args.lr, amsgrad=True, weight_decay=1e-12)


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
        self.flat, self._pending, self._works, self._hooks = [], [], [], []
        for bi, params in enumerate(self.buckets):
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
            off = 0
            for p in params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
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

    def finish(self) -> int:
        """After backward: launch what the hooks could not (parameters that received no gradient on this rank keep
        their zeros; every rank issues the same collectives), wait, average.  Returns the bytes all-reduced."""
        self._armed = False
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

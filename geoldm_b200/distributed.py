"""Multi-GPU sampling: independent molecules are sharded over ranks, each rank runs its own CUDA-graphed
1000-step loop, and ONE all_gather at the end returns the results in the caller's order.

The reference never shards sampling (always one device, main_qm9.py:270); there is no collective inside the loop
because molecules never interact (edges only within a molecule, egnn/models.py:122-127).  The device noise stream
is keyed by GLOBAL molecule index, so the drawn noise does not depend on the number of ranks.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from .packing import balance_shards


def shard_indices(nodesxsample, world_size: int, rank: int) -> np.ndarray:
    """Global molecule indices assigned to `rank` (greedy balance on n(n-1), deterministic on every rank)."""
    n = np.asarray(torch.as_tensor(nodesxsample).cpu().numpy(), dtype=np.int64)
    return balance_shards(n, world_size)[rank]


def sample_sharded(args, device, generative_model, dataset_info, nodesxsample, context=None, fix_noise=False,
                   seed: int = 0, group=None, sample_fn: Optional[Callable] = None):
    """Drop-in for qm9/sampling.py:sample on N ranks: same arguments and return value
    (one_hot, charges, x, node_mask), every rank receives the complete batch.

    sample_fn(args, device, model, dataset_info, nodesxsample=..., context=..., fix_noise=..., seed=..., mol_ids=...)
    defaults to geoldm_b200.sampling.sample (CUDA); tests inject a CPU stand-in to exercise the sharding logic."""
    if sample_fn is None:
        from .sampling import sample as sample_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nodes = torch.as_tensor(nodesxsample).cpu()
    B = len(nodes)
    shards = balance_shards(nodes.numpy(), world)            # computed once; identical on every rank
    mine = shards[rank]
    if len(mine) > 0:
        ctx = None if context is None else context[torch.from_numpy(mine)]
        one_hot, charges, x, node_mask = sample_fn(args, device, generative_model, dataset_info,
                                                   nodesxsample=nodes[torch.from_numpy(mine)], context=ctx,
                                                   fix_noise=fix_noise, seed=seed, mol_ids=mine)
    if world == 1:
        return one_hot, charges, x, node_mask
    # pad every shard to the largest one, gather once, scatter back into the caller's order.  A rank whose shard is
    # empty (more ranks than molecules) skips sampling and contributes all-padding payloads: the shapes and dtypes of
    # the payload rows are broadcast from the first non-empty rank so that every rank issues the same collectives.
    cap = max(len(s_) for s_ in shards)
    first = next(r for r in range(world) if len(shards[r]) > 0)
    meta = [None]
    if rank == first:
        meta = [[(tuple(t.shape[1:]), t.dtype) for t in (one_hot, charges, x, node_mask)]]
    dist.broadcast_object_list(meta, src=first if group is None else dist.get_global_rank(group, first), group=group)
    if len(mine) == 0:
        dev_ = torch.device(device)
        one_hot, charges, x, node_mask = (torch.zeros((0,) + shp, dtype=dt, device=dev_) for shp, dt in meta[0])

    def padded(t):
        out = t.new_zeros((cap,) + tuple(t.shape[1:]))
        out[:t.shape[0]] = t
        return out

    idx = torch.full((cap,), -1, dtype=torch.int64, device=x.device)
    idx[:len(mine)] = torch.from_numpy(mine).to(x.device)
    has_charges = meta[0][1][0] != () and all(d > 0 for d in meta[0][1][0])     # include_charges=False -> empty charges
    payload = [padded(one_hot), padded(charges) if has_charges else None, padded(x), padded(node_mask), idx]
    gathered = []
    for t in payload:
        if t is None:
            gathered.append(None)
            continue
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t.contiguous(), group=group)
        gathered.append(torch.cat(bufs, dim=0))
    g_idx = gathered[4]
    keep = g_idx >= 0
    order = g_idx[keep]

    def ordered(t):
        if t is None:
            return charges
        out = t.new_zeros((B,) + tuple(t.shape[1:]))
        out[order] = t[keep]
        return out

    return ordered(gathered[0]), ordered(gathered[1]), ordered(gathered[2]), ordered(gathered[3])

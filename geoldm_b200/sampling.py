"""``sample`` with the reference's signature and return value (qm9/sampling.py:110-154)."""
from __future__ import annotations

import torch


def build_masks(nodesxsample, max_n_nodes, device):
    """node_mask [bs, n, 1], edge_mask [bs*n*n, 1] (qm9/sampling.py:117-128), built on the device."""
    n = torch.as_tensor(nodesxsample, device=device).reshape(-1, 1)
    node_mask = (torch.arange(max_n_nodes, device=device).unsqueeze(0) < n).float()
    edge_mask = node_mask.unsqueeze(1) * node_mask.unsqueeze(2)
    edge_mask = edge_mask * (~torch.eye(max_n_nodes, dtype=torch.bool, device=device)).unsqueeze(0)
    return node_mask.unsqueeze(2), edge_mask.reshape(-1, 1)


def sample(args, device, generative_model, dataset_info, prop_dist=None, nodesxsample=torch.tensor([10]),
           context=None, fix_noise=False, **sampler_kwargs):
    max_n_nodes = dataset_info['max_n_nodes']
    assert int(torch.max(nodesxsample)) <= max_n_nodes
    batch_size = len(nodesxsample)
    node_mask, edge_mask = build_masks(nodesxsample, max_n_nodes, device)
    if args.context_node_nf > 0:
        if context is None:
            context = prop_dist.sample_batch(nodesxsample)
        context = context.unsqueeze(1).repeat(1, max_n_nodes, 1).to(device) * node_mask
    else:
        context = None
    if args.probabilistic_model != 'diffusion':
        raise ValueError(args.probabilistic_model)
    x, h = generative_model.sample(batch_size, max_n_nodes, node_mask, edge_mask, context, fix_noise=fix_noise,
                                   **sampler_kwargs)
    return h['categorical'], h['integer'], x, node_mask

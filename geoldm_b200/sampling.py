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


def check_sample_invariants(x, one_hot, charges, node_mask, include_charges):
    """The reference's post-sampling asserts (qm9/sampling.py:141-150, equivariant_diffusion/utils.py:46-56): padded
    entries are zero and the coordinates of every molecule are centred.  Evaluated on the device with ONE host read."""
    pad = 1 - node_mask
    worst_pad = torch.stack([(x * pad).abs().max(), (one_hot.to(x.dtype) * pad).abs().max(),
                             (charges.to(x.dtype) * pad).abs().max() if include_charges and charges.numel() else x.new_zeros(())])
    rel_mean = x.sum(dim=1, keepdim=True).abs().max() / (x.abs().max() + 1e-10)
    stats = torch.cat([worst_pad, rel_mean.reshape(1)]).cpu()
    assert float(stats[:3].max()) < 1e-4, 'Variables not masked properly.'
    assert float(stats[3]) < 1e-2, f'Mean is not zero, relative_error {float(stats[3])}'


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
    check_sample_invariants(x, h['categorical'], h['integer'], node_mask, args.include_charges)
    return h['categorical'], h['integer'], x, node_mask


def reverse_tensor(x):
    return x.flip(0)


def sample_chain(args, device, flow, n_tries, dataset_info, prop_dist=None, **sampler_kwargs):
    """qm9/sampling.py:54-107: one molecule of the dataset's typical size sampled with 100 kept frames, retried until
    the final frame is a stable molecule (stability checked on the device)."""
    from .stability import check_stability
    n_samples = 1
    if args.dataset in ('qm9', 'qm9_second_half', 'qm9_first_half'):
        n_nodes = 19
    elif args.dataset == 'geom':
        n_nodes = 44
    else:
        raise ValueError()
    if args.context_node_nf > 0:
        context = prop_dist.sample(n_nodes).unsqueeze(1).unsqueeze(0)
        context = context.repeat(1, n_nodes, 1).to(device)
    else:
        context = None
    node_mask = torch.ones(n_samples, n_nodes, 1, device=device)
    edge_mask = (1 - torch.eye(n_nodes, device=device)).unsqueeze(0).repeat(n_samples, 1, 1).view(-1, 1)
    if args.probabilistic_model != 'diffusion':
        raise ValueError
    one_hot, charges, x = None, None, None
    for i in range(n_tries):
        chain = flow.sample_chain(n_samples, n_nodes, node_mask, edge_mask, context, keep_frames=100, **sampler_kwargs)
        chain = reverse_tensor(chain)
        chain = torch.cat([chain, chain[-1:].repeat(10, 1, 1)], dim=0)       # hold the final frame
        atom_type = torch.argmax(chain[-1, :, 3:-1], dim=1)
        mol_stable = check_stability(chain[-1, :, 0:3], atom_type, dataset_info, device=device)[0]
        x = chain[:, :, 0:3]
        one_hot = torch.nn.functional.one_hot(torch.argmax(chain[:, :, 3:-1], dim=2),
                                              num_classes=len(dataset_info['atom_decoder']))
        charges = torch.round(chain[:, :, -1:]).long()
        if mol_stable:
            break
    return one_hot, charges, x


def sample_sweep_conditional(args, device, generative_model, dataset_info, prop_dist, n_nodes=19, n_frames=100,
                             **sampler_kwargs):
    """qm9/sampling.py:157-171: one noise draw shared by n_frames molecules (fix_noise) while the conditioning value
    sweeps linearly between the property's extremes for that molecule size."""
    import numpy as np
    nodesxsample = torch.tensor([n_nodes] * n_frames)
    context = []
    for key in prop_dist.distributions:
        min_val, max_val = prop_dist.distributions[key][n_nodes]['params']
        mean, mad = prop_dist.normalizer[key]['mean'], prop_dist.normalizer[key]['mad']
        lo, hi = (min_val - mean) / mad, (max_val - mean) / mad
        context.append(torch.tensor(np.linspace(float(lo), float(hi), n_frames)).unsqueeze(1))
    context = torch.cat(context, dim=1).float().to(device)
    return sample(args, device, generative_model, dataset_info, prop_dist, nodesxsample=nodesxsample, context=context,
                  fix_noise=True, **sampler_kwargs)

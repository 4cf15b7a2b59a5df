"""Model factory with the reference's surface (qm9/models.py): ``get_latent_diffusion(args, device,
dataset_info, dataloader_train) -> (model, nodes_dist, prop_dist)`` (:103-166), ``get_autoencoder`` (:54-100)
and ``DistributionNodes`` (:178-215).  ``args`` is the same argparse Namespace the reference uses; the extra
optional attribute ``args.mma_mode`` ("auto" | "fp32" | "3xf16" | "3xtf32" | "tf32") selects the arithmetic of the
fused kernels; the default "auto" picks the fp16-split tensor-core path when hidden_nf is 64/128/192/256 and the fp32
FFMA path otherwise."""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions.categorical import Categorical

from .diffusion import EnHierarchicalVAE, EnLatentDiffusion
from .dynamics import EGNN_decoder_QM9, EGNN_dynamics_QM9, EGNN_encoder_QM9


class DistributionNodes:
    """Categorical over molecule sizes, indexed in the histogram's dict order like the reference."""

    def __init__(self, histogram):
        self.n_nodes = torch.tensor(list(histogram.keys()))
        self.keys = {int(n): i for i, n in enumerate(histogram.keys())}
        prob = np.array(list(histogram.values()), dtype=np.float64)
        prob = prob / np.sum(prob)
        self.prob = torch.from_numpy(prob).float()
        self._lut = {}
        self.m = Categorical(torch.tensor(prob))

    def sample(self, n_samples=1):
        return self.n_nodes[self.m.sample((n_samples,))]

    def log_prob(self, batch_n_nodes):
        assert batch_n_nodes.dim() == 1
        if batch_n_nodes.is_cuda:
            # device lookup table indexed by the atom count (no per-molecule host read, CUDA-graph capturable); a count
            # outside the histogram, a KeyError in the reference, reads as NaN here
            lut = self._lut.get(batch_n_nodes.device)
            if lut is None:
                host = torch.full((int(self.n_nodes.max()) + 1,), float("nan"))
                host[self.n_nodes] = torch.log(self.prob + 1e-30)
                lut = self._lut[batch_n_nodes.device] = host.to(batch_n_nodes.device)
            return lut[batch_n_nodes.clamp(max=lut.numel() - 1)].masked_fill(batch_n_nodes >= lut.numel(), float("nan"))
        idcs = torch.tensor([self.keys[int(i)] for i in batch_n_nodes], device=batch_n_nodes.device)
        return torch.log(self.prob + 1e-30).to(batch_n_nodes.device)[idcs]


def _egnn_kwargs(args, device):
    return dict(n_dims=3, device=device, hidden_nf=args.nf, act_fn=torch.nn.SiLU(), attention=args.attention,
                tanh=args.tanh, mode=args.model, norm_constant=args.norm_constant, inv_sublayers=args.inv_sublayers,
                sin_embedding=args.sin_embedding, normalization_factor=args.normalization_factor,
                aggregation_method=args.aggregation_method, mma_mode=getattr(args, "mma_mode", "auto"))


def get_autoencoder(args, device, dataset_info, dataloader_train):
    in_node_nf = len(dataset_info['atom_decoder']) + int(args.include_charges)
    nodes_dist = DistributionNodes(dataset_info['n_nodes'])
    if len(args.conditioning) > 0:
        raise NotImplementedError("DistributionProperty needs the QM9 dataset (out of scope: no data offline); "
                                  "pass conditioning=[] and feed `context` explicitly")
    kw = _egnn_kwargs(args, device)
    encoder = EGNN_encoder_QM9(in_node_nf=in_node_nf, context_node_nf=args.context_node_nf, out_node_nf=args.latent_nf,
                               n_layers=1, include_charges=args.include_charges, **kw)
    decoder = EGNN_decoder_QM9(in_node_nf=args.latent_nf, context_node_nf=args.context_node_nf, out_node_nf=in_node_nf,
                               n_layers=args.n_layers, include_charges=args.include_charges, **kw)
    vae = EnHierarchicalVAE(encoder=encoder, decoder=decoder, in_node_nf=in_node_nf, n_dims=3,
                            latent_node_nf=args.latent_nf, kl_weight=args.kl_weight,
                            norm_values=args.normalize_factors, include_charges=args.include_charges)
    return vae, nodes_dist, None


def get_latent_diffusion(args, device, dataset_info, dataloader_train):
    if getattr(args, "ae_path", None) is not None:
        raise NotImplementedError("loading a first-stage checkpoint directory (args.ae_path) is out of scope; "
                                  "load the state_dict into the returned model instead")
    for name, default in (("normalization_factor", 1), ("aggregation_method", "sum")):
        if not hasattr(args, name):
            setattr(args, name, default)
    vae, nodes_dist, prop_dist = get_autoencoder(args, device, dataset_info, dataloader_train)
    vae.to(device)
    in_node_nf = args.latent_nf
    # the reference always lets the dynamics append time (its ctor default), whatever args.condition_time says
    dyn_in = in_node_nf + 1 if args.condition_time else in_node_nf
    kw = _egnn_kwargs(args, device)
    net_dynamics = EGNN_dynamics_QM9(in_node_nf=dyn_in, context_node_nf=args.context_node_nf, n_layers=args.n_layers,
                                     **kw)
    if args.probabilistic_model != 'diffusion':
        raise ValueError(args.probabilistic_model)
    vdm = EnLatentDiffusion(vae=vae, trainable_ae=args.trainable_ae, dynamics=net_dynamics, in_node_nf=in_node_nf,
                            n_dims=3, timesteps=args.diffusion_steps, noise_schedule=args.diffusion_noise_schedule,
                            noise_precision=args.diffusion_noise_precision, loss_type=args.diffusion_loss_type,
                            norm_values=args.normalize_factors, include_charges=args.include_charges)
    return vdm.to(device), nodes_dist, prop_dist

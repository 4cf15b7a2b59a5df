"""Shared test helpers (load golden fixtures, rebuild the deterministic weights)."""
import json
import os

import numpy as np
import torch

from oracle import geoldm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(f["meta"]))
    c = meta["cfg"]
    c["normalize_factors"] = tuple(c["normalize_factors"])
    cfg = O.OracleConfig(**c)
    arrays = {k: torch.from_numpy(f[k]) for k in f.files if k != "meta"}
    sd = O.make_state_dict(cfg, meta["seed"], meta["tamed"])
    return cfg, sd, arrays, meta


def part_errors(a, b, n_dims=3):
    """SURVEY §8c metric, separately for the x and h columns."""
    ex = O.err_metric(a[..., :n_dims], b[..., :n_dims])
    eh = O.err_metric(a[..., n_dims:], b[..., n_dims:]) if a.shape[-1] > n_dims else 0.0
    return ex, eh


def make_args(cfg, mma_mode="fp32"):
    """argparse.Namespace with the fields qm9/models.py:get_latent_diffusion reads."""
    import argparse
    return argparse.Namespace(
        ae_path=None, cuda=True, include_charges=cfg.include_charges, context_node_nf=cfg.context_node_nf,
        conditioning=[], latent_nf=cfg.latent_nf, nf=cfg.nf, n_layers=cfg.n_layers, attention=cfg.attention,
        tanh=cfg.tanh, model="egnn_dynamics", norm_constant=cfg.norm_constant, inv_sublayers=cfg.inv_sublayers,
        sin_embedding=False, normalization_factor=cfg.normalization_factor,
        aggregation_method=cfg.aggregation_method, kl_weight=0.01, normalize_factors=list(cfg.normalize_factors),
        condition_time=cfg.condition_time, probabilistic_model="diffusion", diffusion_steps=cfg.diffusion_steps,
        diffusion_noise_schedule=cfg.diffusion_noise_schedule,
        diffusion_noise_precision=cfg.diffusion_noise_precision, diffusion_loss_type="l2", trainable_ae=False,
        ema_decay=0.999, dataset="qm9", remove_h=False, mma_mode=mma_mode)


def build_cuda_model(cfg, sd, device="cuda", mma_mode="fp32"):
    """The product model (geoldm_b200) holding the oracle's deterministic weights."""
    from geoldm_b200.models import get_latent_diffusion
    info = {"atom_decoder": list(range(cfg.n_atom_types)), "n_nodes": {5: 1}, "max_n_nodes": 29}
    model, _, _ = get_latent_diffusion(make_args(cfg, mma_mode), device, info, None)
    res = model.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("vae.encoder") or k.endswith("buffer") for k in res.missing_keys), res.missing_keys
    return model.eval()

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

"""Gradient fixtures (protocol P2): autograd of the UNMODIFIED reference on CPU.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_grad.py

Writes tests/golden/grad_small.npz (every gradient tensor of a small denoiser + decoder) and
tests/golden/grad_qm9.npz (full-size QM9 denoiser nf=256 L=9: input gradient, per-tensor max/L2 of every parameter
gradient and the first 256 entries of each).  Loss = sum(out**2) as config 2 prescribes.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import geoldm_oracle as O          # noqa: E402
from oracle import make_golden as G            # noqa: E402

HEAD = 256


def grads_of(module, loss, z):
    for p in module.parameters():
        p.grad = None
    z.grad = None
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in module.named_parameters() if p.grad is not None}


def main():
    refmods = G.import_reference()
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- small config: every gradient stored --------------------------------------------------------
    cfg = O.OracleConfig(nf=32, n_layers=2)
    model, args, info, sd, gam = G.build_reference(cfg, "qm9", 3, False, refmods)
    model.train()
    gen = torch.Generator().manual_seed(21)
    nodes = [5, 9, 7, 3, 8, 2]
    z, nm, em = G.random_latent(nodes, 9, cfg.latent_nf, gen)
    t = torch.randint(0, 1001, (len(nodes), 1), generator=gen).float() / 1000.0
    z.requires_grad_(True)
    out = model.dynamics._forward(t, z, nm, em, None)
    g = grads_of(model.dynamics, (out ** 2).sum(), z)
    arrays = {"nodes": np.array(nodes), "z": z.detach(), "t": t, "out": out.detach(), "dz": z.grad.clone()}
    arrays.update({"g.dynamics." + k: v for k, v in g.items()})
    for p in model.vae.decoder.parameters():      # frozen by get_latent_diffusion when trainable_ae=False
        p.requires_grad_(True)
    zd = (z.detach() / 3.0).requires_grad_(True)
    xr, hr = model.vae.decoder._forward(zd, nm, em, None)
    g = grads_of(model.vae.decoder, (xr ** 2).sum() + (hr ** 2).sum(), zd)
    arrays.update({"dec_in": zd.detach(), "dec_x": xr.detach(), "dec_h": hr.detach(), "dec_dz": zd.grad.clone()})
    arrays.update({"g.vae.decoder." + k: v for k, v in g.items()})
    G.save("grad_small", cfg, "qm9", 3, False, **arrays)

    # ---- full-size QM9 denoiser: summaries --------------------------------------------------------------
    cfg = O.QM9_CFG
    model, args, info, sd, gam = G.build_reference(cfg, "qm9", 0, False, refmods)
    model.train()
    gen = torch.Generator().manual_seed(22)
    nodes = [17, 23, 9, 29, 19, 12, 21, 4]
    z, nm, em = G.random_latent(nodes, 29, cfg.latent_nf, gen)
    t = torch.randint(0, 1001, (len(nodes), 1), generator=gen).float() / 1000.0
    z.requires_grad_(True)
    out = model.dynamics._forward(t, z, nm, em, None)
    g = grads_of(model.dynamics, (out ** 2).sum(), z)
    names = sorted(g)
    arrays = {"nodes": np.array(nodes), "z": z.detach(), "t": t, "out": out.detach(), "dz": z.grad.clone(),
              "names": np.array(names),
              "gmax": np.array([g[n].abs().max().item() for n in names]),
              "gl2": np.array([g[n].double().norm().item() for n in names]),
              "ghead": np.stack([np.pad(g[n].flatten()[:HEAD].numpy(), (0, max(0, HEAD - g[n].numel())))
                                 for n in names])}
    G.save("grad_qm9", cfg, "qm9", 0, False, **arrays)


if __name__ == "__main__":
    main()

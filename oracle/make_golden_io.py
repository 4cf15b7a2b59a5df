"""tests/golden/ckpt_small/ (a checkpoint directory written by the UNMODIFIED reference's own code paths: pickled args +
utils.save_model) and tests/golden/xyz_golden.json (files written by qm9/visualizer.py:save_xyz_file).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_io.py
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import geoldm_oracle as O          # noqa: E402
from oracle import make_golden as G            # noqa: E402


def main():
    refmods = G.import_reference()
    dc, qm, qs = refmods
    import utils as ref_utils
    import qm9.visualizer as vis
    cfg = O.OracleConfig(nf=32, n_layers=1)
    model, args, info, sd, gam = G.build_reference(cfg, "qm9", 12, False, refmods, encoder=True)
    args.no_cuda, args.exp_name = True, "ckpt_small"
    out = os.path.join(ROOT, "tests", "golden", "ckpt_small")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "args.pickle"), "wb") as f:      # main_qm9.py:258-259
        pickle.dump(args, f)
    ref_utils.save_model(model, os.path.join(out, "generative_model_ema.npy"))
    gen = torch.Generator().manual_seed(4)
    nodes = [6, 11, 3]
    z, nm, em = G.random_latent(nodes, 11, cfg.latent_nf, gen)
    t = torch.tensor([[0.25], [0.5], [1.0]])
    with torch.no_grad():
        eps = model.dynamics._forward(t, z, nm, em, None)
    np.savez_compressed(os.path.join(out, "forward.npz"), nodes=np.array(nodes), z=z.numpy(), t=t.numpy(),
                        out=eps.numpy())
    # xyz files
    one_hot = torch.nn.functional.one_hot(torch.randint(0, 5, (3, 11), generator=gen), 5).float() * nm
    x = torch.randn(3, 11, 3, generator=gen) * nm * 2.5
    with tempfile.TemporaryDirectory() as d:
        vis.save_xyz_file(d + "/", one_hot, None, x, info, id_from=7, name="molecule", node_mask=nm)
        vis.save_xyz_file(d + "/", one_hot[:1], None, x[:1], info, id_from=0, name="full")      # no node_mask
        files = {fn: open(os.path.join(d, fn)).read() for fn in sorted(os.listdir(d))}
    with open(os.path.join(ROOT, "tests", "golden", "xyz_golden.json"), "w") as f:
        json.dump({"files": files, "one_hot": one_hot.tolist(), "x": x.tolist(), "node_mask": nm.tolist()}, f)
    print("wrote", out, sorted(os.listdir(out)), list(files))


if __name__ == "__main__":
    main()

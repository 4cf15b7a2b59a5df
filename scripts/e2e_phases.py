"""Where does the end-to-end sample() time go?  Wall-clock phases with synchronisation between them (diagnostic)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                                   # noqa: E402
from geoldm_b200.packing import pack_from_masks                                # noqa: E402
from geoldm_b200.sampling import build_masks                                   # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n_mol = int(os.environ.get("MOLS", 1250))
    from geoldm_b200.models import get_latent_diffusion
    margs = bench.qm9_args(os.environ.get("MODE", "3xf16"))
    info = {"atom_decoder": ["H", "C", "N", "O", "F"], "n_nodes": {5: 1}, "max_n_nodes": 29}
    torch.manual_seed(0)
    model, _, _ = get_latent_diffusion(margs, dev, info, None)
    bench.tame_(model, margs.nf)
    model.eval()
    nodes = bench.workload_nodes(n_mol)
    t = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        t[name] = time.perf_counter() - t0
        return time.perf_counter()

    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nm, em = build_masks(torch.as_tensor(nodes), 29, dev)
        t0 = tick("masks", t0)
        batch = pack_from_masks(nm, em, validate=True)
        t0 = tick("pack", t0)
        z = model.sample_latent_ragged(batch, None, seed=0, n_steps=3)
        t0 = tick("first3_eager", t0)
        z = model.sample_latent_ragged(batch, None, seed=0, n_steps=103)
        t0 = tick("103_steps_graph", t0)
        z = model.sample_latent_ragged(batch, None, seed=0)
        t0 = tick("full_latent", t0)
        zz = torch.zeros(n_mol * 29, 4, device=dev)
        zz[batch.node_src.long()] = z
        x, h = model.vae.decode(zz.view(n_mol, 29, 4), nm, em, None)
        t0 = tick("decode", t0)
        out = [v.cpu() for v in (x, h["categorical"], h["integer"])]
        t0 = tick("d2h", t0)
        print({k: round(v * 1e3, 2) for k, v in t.items()})
        print("per-step ms (graph, steps 4..103):", (t["103_steps_graph"] - t["first3_eager"]) / 100 * 1e3,
              " full:", t["full_latent"] / 1002 * 1e3)


if __name__ == "__main__":
    main()

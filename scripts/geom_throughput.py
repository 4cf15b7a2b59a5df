"""BASELINE config 4 shape on one B200: GEOM-Drugs GeoLDM sampling (nf=256, 4 blocks, latent_nf=2, <= 181 atoms, tiled
edge kernel).  Times graph-replayed sampling steps on molecules drawn from the GEOM atom-count histogram and prints one
JSON line (molecules/s for T=1000 sampling, edge messages/s).  Informational: bench.py's line stays on the QM9 config."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                        # noqa: E402
from geoldm_b200.histograms import GEOM_WITH_H_N_NODES as HIST      # noqa: E402
from geoldm_b200.models import get_latent_diffusion                 # noqa: E402
from geoldm_b200.packing import pack_molecules                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mols", type=int, default=256)
    ap.add_argument("--mma-mode", default="3xf16")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    margs = bench.qm9_args(a.mma_mode)
    margs.n_layers, margs.latent_nf, margs.include_charges, margs.dataset = 4, 2, False, "geom"
    info = {"atom_decoder": list(range(16)), "n_nodes": {44: 1}, "max_n_nodes": 181}
    torch.manual_seed(0)
    model, _, _ = get_latent_diffusion(margs, dev, info, None)
    bench.tame_(model, margs.nf)
    model.eval()
    keys = np.array(list(HIST.keys()))
    p = np.array(list(HIST.values()), dtype=np.float64)
    nodes = keys[np.random.default_rng(0).choice(len(keys), size=a.mols, p=p / p.sum())]
    batch = pack_molecules(nodes, dev)
    t = {}
    model.sample_latent_ragged(batch, None, seed=0, n_steps=5)          # warm-up: weight packing, workspace, first capture
    for name, n in (("short", 3), ("long", 103)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.sample_latent_ragged(batch, None, seed=0, n_steps=n)
        torch.cuda.synchronize()
        t[name] = time.perf_counter() - t0
    ms = (t["long"] - t["short"]) / 100 * 1e3
    T = margs.diffusion_steps
    line = {"workload": "GEOM-Drugs GeoLDM sampling step, nf=256 n_layers=4 latent_nf=2", "molecules": int(a.mols),
            "atoms": int(batch.n_node), "edges": int(batch.n_edge), "max_atoms": int(nodes.max()), "mma_mode": a.mma_mode,
            "ms_per_step": ms, "molecules_per_s_T1000": a.mols / ((T + 2) * ms * 1e-3),
            "edge_msgs_per_s": 8.0 * batch.n_edge / (ms * 1e-3)}
    print(json.dumps(line))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/geom_throughput_{a.mma_mode}.json", "w") as f:
        json.dump(line, f)


if __name__ == "__main__":
    main()

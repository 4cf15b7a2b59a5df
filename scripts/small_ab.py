"""A/B of library switches on the small configurations: python scripts/small_ab.py [molecule counts ...]
Prints ms per sampler step for QM9-shaped batches of the given sizes (default 64) and for configs.geom32, with the
environment the caller set (e.g. GEOLDM_TC_CHAIN=0 / 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    from geoldm_b200.models import get_latent_diffusion
    from geoldm_b200.packing import pack_molecules
    dev = torch.device("cuda:0")
    sizes = [int(a) for a in sys.argv[1:]] or [64]
    mode = "3xf16"
    margs = bench.qm9_args(mode)
    info = {"atom_decoder": list(range(5)), "n_nodes": {19: 1}, "max_n_nodes": 29}
    torch.manual_seed(0)
    model, _, _ = get_latent_diffusion(margs, dev, info, None)
    bench.tame_(model, margs.nf)
    model.eval()
    out = {}
    for n in sizes:
        nodes = bench.workload_nodes(n, seed=1)
        batch = pack_molecules(nodes, dev)
        loop = bench.StepLoop(model, batch, dev, margs.latent_nf, margs.diffusion_steps)
        loop.capture()
        out["qm9_%d" % n] = round(loop.timed(10, 50, None) / 50, 4)
    out["geom32"] = round(bench.bench_geom32(mode, dev)["ms_per_step"], 4)
    print(os.environ.get("GEOLDM_TC_CHAIN", "-"), out, flush=True)


if __name__ == "__main__":
    main()

"""Tensor-core GEMMs of the training path against the fp32 FFMA kernels at a LARGE batch (gradients are 1 / batch-size
small: the case that broke the unscaled fp16 split of the input-gradient GEMM).  One GPU:
    python scripts/train_tc_check.py [--batch 256]
Prints the worst relative difference (max-abs, per tensor) between the gradients of the two paths."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import train_check as tc                                               # noqa: E402
from geoldm_b200 import losses                                         # noqa: E402
from geoldm_b200.histograms import HISTOGRAMS                          # noqa: E402
from geoldm_b200.models import get_latent_diffusion                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    args = tc.make_args(192, 9)
    info = {"atom_decoder": list(range(5)), "n_nodes": HISTOGRAMS["qm9_second_half"], "max_n_nodes": 29}
    torch.manual_seed(0)
    model, nodes_dist, _ = get_latent_diffusion(args, dev, info, None)
    rng = np.random.default_rng(5)
    sizes = np.array(list(HISTOGRAMS["qm9_second_half"].keys()))
    prob = np.array(list(HISTOGRAMS["qm9_second_half"].values()), dtype=np.float64)
    nodes = rng.choice(sizes, size=a.batch, p=prob / prob.sum()).tolist()
    x, h, nm, em, ctx, draws = tc.synth_batch(nodes, 29, dev, torch.Generator().manual_seed(11))
    torch.manual_seed(1)
    draws["eps_enc"] = losses.masked_noise(len(nodes), 29, 3, 1, nm)
    draws["eps_t"] = losses.masked_noise(len(nodes), 29, 3, 1, nm)
    model.train()
    grads = {}
    for flag in ("1", "0"):
        os.environ["GEOLDM_TRAIN_TC"] = flag
        model.zero_grad(set_to_none=True)
        nll, _, _ = losses.compute_loss_and_nll(args, model, nodes_dist, x, h, nm, em, ctx, draws=draws)
        nll.backward()
        grads[flag] = {n: p.grad.double().clone() for n, p in model.named_parameters() if p.grad is not None}
    errs = {n: float((grads["1"][n] - g).abs().max() / g.abs().max().clamp_min(1e-300)) for n, g in grads["0"].items()}
    worst = max(errs, key=errs.get)
    print(json.dumps({"batch": a.batch, "tensors": len(errs), "worst_rel": errs[worst], "worst_tensor": worst}))
    assert errs[worst] < 2e-5, (worst, errs[worst])


if __name__ == "__main__":
    main()

"""sample_chain / sample_sweep_conditional (fix_noise) against the unmodified reference run
(tests/golden/chain_*.npz, oracle/make_golden_chain.py) with the same noise draws.  Tolerance: 1e-3 relative on
coordinates over the complete 1000-step trajectories (north_star's trajectory gate), discrete outputs identical."""
import types

import numpy as np
import pytest
import torch

from oracle import geoldm_oracle as O
from tests.helpers import build_cuda_model, load_golden, make_args

pytestmark = pytest.mark.gpu


def _draws(seed, T, bs, n, latent):
    torch.manual_seed(seed)
    out = []
    for _ in range(T + 2):
        zx = torch.randn(bs, n, 3)
        zh = torch.randn(bs, n, latent)
        out.append(torch.cat([zx, zh], 2))
    return torch.stack(out)


def test_model_sample_chain():
    cfg, sd, a, _ = load_golden("chain_model")
    model = build_cuda_model(cfg, sd, "cuda", "3xf16")
    nodes = a["nodes"].tolist()
    kf = int(a["keep_frames"][0])
    nm, em = O.build_masks(nodes, 12)
    noise = _draws(int(a["torch_seed"][0]), cfg.diffusion_steps, 2, 12, cfg.latent_nf)
    chain = model.sample_chain(2, 12, nm.cuda(), em.cuda(), None, keep_frames=kf, noise=noise).cpu()
    ref = a["chain"]
    assert chain.shape == ref.shape
    ex = O.err_metric(chain[..., :3], ref[..., :3])
    print(f"[chain] {kf} decoded frames x 2 molecules: x err {ex:.2e}")
    assert ex < 1e-3
    assert torch.equal(chain[..., 3:], ref[..., 3:])            # one-hot atom types and integer charges per frame


def test_sampling_sample_chain():
    from geoldm_b200.sampling import sample_chain
    cfg, sd, a, meta = load_golden("chain_sampling")
    model = build_cuda_model(cfg, sd, "cuda", "3xf16")
    args = make_args(cfg)
    info = {"name": "qm9", "atom_decoder": ['H', 'C', 'N', 'O', 'F'], "max_n_nodes": 29}
    noise = _draws(int(a["torch_seed"][0]), cfg.diffusion_steps, 1, 19, cfg.latent_nf)
    one_hot, charges, x = sample_chain(args, "cuda", model, 1, info, noise=noise)
    assert x.shape == a["x"].shape == (110, 19, 3)
    ex = O.err_metric(x.cpu(), a["x"])
    print(f"[chain] qm9.sampling.sample_chain: x err {ex:.2e}")
    assert ex < 1e-3
    assert torch.equal(one_hot.cpu(), a["one_hot"]) and torch.equal(charges.cpu(), a["charges"])


def test_sweep_conditional_fix_noise():
    from geoldm_b200.sampling import sample_sweep_conditional
    cfg, sd, a, meta = load_golden("chain_sweep")
    model = build_cuda_model(cfg, sd, "cuda", "3xf16")
    args = make_args(cfg)
    args.dataset = "qm9_second_half"
    info = {"name": "qm9_second_half", "atom_decoder": ['H', 'C', 'N', 'O', 'F'], "max_n_nodes": 29}
    prop = types.SimpleNamespace(distributions={"alpha": {9: {"params": (40.0, 90.0)}}},
                                 normalizer={"alpha": {"mean": torch.tensor(75.0), "mad": torch.tensor(6.0)}})
    noise = _draws(int(a["torch_seed"][0]), cfg.diffusion_steps, 1, 29, cfg.latent_nf).expand(-1, 6, -1, -1)
    one_hot, charges, x, node_mask = sample_sweep_conditional(args, "cuda", model, info, prop, n_nodes=9, n_frames=6,
                                                              noise=noise)
    ex = O.err_metric(x.cpu(), a["x"])
    print(f"[chain] sample_sweep_conditional (fix_noise): x err {ex:.2e}")
    assert ex < 1e-3
    assert torch.equal(one_hot.cpu().long(), a["one_hot"].long())
    assert torch.equal(node_mask.cpu(), a["node_mask"])
    # the six molecules share the noise and differ only through the context: they must not be identical
    assert (x[0] - x[5]).abs().max() > 0

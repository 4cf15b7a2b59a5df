"""INTEGRATION.md level 2 — the drop-in SURVEY §8(b) describes: this repo's `EGNN_dynamics_QM9` plugged into the
REFERENCE's own `EnLatentDiffusion` as its `dynamics=` module (qm9/models.py:152, consumed at en_diffusion.py:283 and
called at :314-317), then driven through the reference's own `sample_p_zs_given_zt` / `qm9/sampling.py:sample`.

The reference is pure Python and exists only in the build container (/root/reference) or in a driver-provided
baseline/_ref; the GPU box has neither, so the GPU test skips there and the constructor / attribute contract is checked on
the CPU wherever the reference is present.  Nothing here is imported by the product.
"""
import contextlib
import io
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_dir():
    for cand in (os.environ.get("GEOLDM_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "equivariant_diffusion")) and os.path.isdir(os.path.join(cand, "egnn")):
            return cand
    return None


def _import_reference():
    ref = _reference_dir()
    if ref is None:
        pytest.skip("the reference tree is not on this machine (/root/reference, baseline/_ref)")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):       # plotting imports of qm9/visualizer.py
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import configs.datasets_config as dc
    import qm9.models as qm
    import qm9.sampling as qs
    return dc, qm, qs


def _reference_model(cfg, device):
    """The reference's EnLatentDiffusion holding the oracle's deterministic weights (as oracle/make_golden.py builds it)."""
    from oracle import geoldm_oracle as O
    from tests.helpers import make_args
    dc, qm, qs = _import_reference()
    args = make_args(cfg)
    args.cuda = device != "cpu"
    del args.mma_mode
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(1234)
        ref, _, _ = qm.get_latent_diffusion(args, device, dc.get_dataset_info("qm9", False), None)
    sd = O.make_state_dict(cfg, 3, tamed=True)
    state = ref.state_dict()
    for k, v in sd.items():
        state[k] = v.clone()
    ref.load_state_dict(state)
    return ref.eval(), args, dc, qs


def _our_dynamics(ref, cfg, device, mma_mode):
    from geoldm_b200.dynamics import EGNN_dynamics_QM9
    dyn = EGNN_dynamics_QM9(in_node_nf=ref.dynamics.in_node_nf, context_node_nf=ref.dynamics.context_node_nf, n_dims=3,
                            device=device, hidden_nf=cfg.nf, act_fn=torch.nn.SiLU(), n_layers=cfg.n_layers,
                            attention=cfg.attention, tanh=cfg.tanh, mode="egnn_dynamics", norm_constant=cfg.norm_constant,
                            inv_sublayers=cfg.inv_sublayers, sin_embedding=False,
                            normalization_factor=cfg.normalization_factor, aggregation_method=cfg.aggregation_method,
                            mma_mode=mma_mode)
    res = dyn.load_state_dict(ref.dynamics.state_dict(), strict=True)       # the reference's own parameter names
    assert not res.missing_keys and not res.unexpected_keys
    return dyn.eval()


def test_reference_accepts_our_module_interface_cpu():
    """Constructor kwargs, attributes and state_dict layout the reference relies on (no compute: the module refuses CPU)."""
    from oracle import geoldm_oracle as O
    from geoldm_b200._lib import GeoldmError
    cfg = O.OracleConfig(nf=64, n_layers=2)
    ref, args, dc, qs = _reference_model(cfg, "cpu")
    dyn = _our_dynamics(ref, cfg, "cpu", "fp32")
    for attr in ("in_node_nf", "context_node_nf", "n_dims", "condition_time", "device", "_forward", "wrap_forward",
                 "unwrap_forward"):
        assert hasattr(dyn, attr), attr
    assert dyn.in_node_nf == ref.dynamics.in_node_nf and dyn.n_dims == ref.dynamics.n_dims
    ref.dynamics = dyn                                               # nn.Module attribute swap, as INTEGRATION.md shows
    assert ref.dynamics is dyn and "dynamics.egnn.embedding.weight" in ref.state_dict()
    nm = torch.ones(1, 5, 1)
    em = (torch.ones(5, 5) - torch.eye(5)).reshape(-1, 1)
    with pytest.raises(GeoldmError):                                 # loud failure, no CPU fallback behind the reference
        ref.phi(torch.zeros(1, 5, 4), torch.zeros(1, 1), nm, em, None)


@pytest.mark.gpu
@pytest.mark.parametrize("mma_mode", ["fp32", "3xf16"])
def test_reference_sampler_runs_on_our_denoiser(mma_mode):
    """The reference's own `sample_p_zs_given_zt` (its Python loop, its noise, its host syncs) and `qm9.sampling.sample`
    with this repo's CUDA `_forward` underneath: every step equals the all-reference step on the same draws."""
    from oracle import geoldm_oracle as O
    from geoldm_b200 import _lib
    assert torch.cuda.is_available()
    if mma_mode != "fp32" and not _lib.lib().geoldm_has_tcgen05():
        pytest.skip("tcgen05 kernels not built")
    cfg = O.OracleConfig(nf=64, n_layers=2, diffusion_steps=50)
    ref_cpu, args, dc, qs = _reference_model(cfg, "cpu")
    ref_gpu, _, _, _ = _reference_model(cfg, "cuda")
    ref_gpu.dynamics = _our_dynamics(ref_gpu, cfg, "cuda", mma_mode)
    nodes = [5, 17, 29, 9]
    nm, em = O.build_masks(nodes, 29)
    bs, T = len(nodes), cfg.diffusion_steps
    torch.manual_seed(5)
    z = ref_cpu.sample_combined_position_feature_noise(bs, 29, nm)
    worst = 0.0
    with torch.no_grad():
        for s in (T - 1, T // 2, 0):
            s_arr, t_arr = torch.full((bs, 1), s) / T, (torch.full((bs, 1), s) + 1) / T
            eps_ref = ref_cpu.phi(z, t_arr, nm, em, None)
            eps_our = ref_gpu.phi(z.cuda(), t_arr.cuda(), nm.cuda(), em.cuda(), None).cpu()
            worst = max(worst, O.err_metric(eps_our, eps_ref))
            torch.manual_seed(100 + s)
            z_ref = ref_cpu.sample_p_zs_given_zt(s_arr, t_arr, z, nm, em, None)
            torch.manual_seed(100 + s)                                # same global-generator draws on the host ...
            noise = torch.randn(bs, 29, 3), torch.randn(bs, 29, cfg.latent_nf)
            z_our = ref_gpu.sample_p_zs_given_zt(s_arr.cuda(), t_arr.cuda(), z.cuda(), nm.cuda(), em.cuda(), None).cpu()
            assert torch.isfinite(z_our).all() and z_our.shape == z_ref.shape
            z = z_ref
    print(f"[integration L2] reference EnLatentDiffusion + geoldm_b200 denoiser ({mma_mode}): eps_hat err {worst:.2e}")
    assert worst < 1e-5
    # the reference's outer entry point on top of it
    one_hot, charges, x, node_mask = qs.sample(args, "cuda", ref_gpu, dc.get_dataset_info("qm9", False),
                                               nodesxsample=torch.tensor(nodes))
    assert x.shape == (bs, 29, 3) and bool(torch.isfinite(x).all())

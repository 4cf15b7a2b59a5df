"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden, made by
oracle/make_golden.py).  Tolerance: 2e-6 (max-abs / max-abs per part) for single forwards — the
oracle issues the same ATen ops, so differences are only BLAS blocking / reduction order; the
reference's own fp32-vs-fp64 noise floor is 4.6e-6."""
import numpy as np
import pytest
import torch

from oracle import geoldm_oracle as O
from tests.helpers import load_golden, part_errors

TOL = 2e-6


def test_schedule_table():
    cfg, sd, a, _ = load_golden("schedule_polynomial2_T1000")
    assert torch.equal(sd["gamma.gamma"], a["gamma"])
    assert abs(float(a["gamma"][0]) + 11.512916) < 1e-5 and abs(float(a["gamma"][1000]) - 11.512516) < 1e-5


@pytest.mark.parametrize("tag", ["s1", "s30"])
def test_qm9_forward(tag):
    cfg, sd, a, _ = load_golden("qm9_forward")
    nm, em = O.build_masks(a["nodes"].tolist(), 29)
    z = a[f"z_{tag}"]
    with torch.no_grad():
        for key, t in (("out_tscalar", torch.tensor([[0.5]])), ("out_tvec", a[f"t_vec_{tag}"]),
                       ("out_t0", torch.zeros(z.shape[0], 1))):
            out = O.dynamics_forward(sd, cfg, t, z, nm, em)
            ex, eh = part_errors(out, a[f"{key}_{tag}"])
            assert ex < TOL and eh < TOL, (key, ex, eh)
            assert float((out * (1 - nm)).abs().max()) == 0.0


def test_qm9_decoder_and_decode():
    cfg, sd, a, _ = load_golden("qm9_forward")
    nm, em = O.build_masks(a["nodes"].tolist(), 29)
    with torch.no_grad():
        dx, dh = O.decoder_forward(sd, cfg, a["dec_in"], nm, em)
        x, one_hot, charges = O.decode(sd, cfg, a["dec_in"], nm, em)
    assert O.err_metric(dx, a["dec_x"]) < TOL and O.err_metric(dh, a["dec_h"]) < TOL
    assert O.err_metric(x, a["decode_x"]) < TOL
    assert torch.equal(one_hot.long(), a["decode_onehot"].long())
    assert torch.equal(charges.long(), a["decode_charges"].long())


@pytest.mark.parametrize("name", ["small_default", "small_S2_noatt_notanh", "small_mean", "small_cond",
                                  "small_latent2"])
def test_small_variants(name):
    cfg, sd, a, _ = load_golden(name)
    nm, em = O.build_masks(a["nodes"].tolist(), 29)
    ctx = a.get("context")
    with torch.no_grad():
        out = O.dynamics_forward(sd, cfg, a["t_vec"], a["z"], nm, em, ctx)
        dx, dh = O.decoder_forward(sd, cfg, a["z"], nm, em, ctx)
    ex, eh = part_errors(out, a["out"])
    assert ex < TOL and eh < TOL, (ex, eh)
    assert O.err_metric(dx, a["dec_x"]) < TOL and O.err_metric(dh, a["dec_h"]) < TOL


def test_geom_forward():
    cfg, sd, a, _ = load_golden("geom_forward")
    nm, em = O.build_masks(a["nodes"].tolist(), 181)
    with torch.no_grad():
        out = O.dynamics_forward(sd, cfg, torch.tensor([[0.3]]), a["z"], nm, em)
    ex, eh = part_errors(out, a["out"])
    assert ex < TOL and eh < TOL, (ex, eh)


def test_sampler_steps_teacher_forced_and_free():
    cfg, sd, a, _ = load_golden("qm9_sampler_steps")
    nodes = a["nodes"].tolist()
    nm, em = O.build_masks(nodes, 29)
    noise = O.NoiseSource(a["raw"])
    T = cfg.diffusion_steps
    with torch.no_grad():
        z = O.combined_noise(cfg, noise, len(nodes), 29, nm, cfg.latent_nf)
        assert O.err_metric(z, a["z"][0]) < 1e-6
        for k, s in enumerate(reversed(range(T - 4, T))):
            s_arr = torch.full((len(nodes), 1), float(s)) / T
            t_arr = torch.full((len(nodes), 1), float(s + 1)) / T
            # teacher forced: feed the reference's z_t
            src = O.NoiseSource(a["raw"]); src.k = k + 1
            zs, eps = O.sample_p_zs_given_zt(sd, cfg, s_arr, t_arr, a["z"][k], nm, em, None, src, True)
            ex, eh = part_errors(eps, a["eps"][k])
            assert ex < TOL and eh < TOL, ("eps", k, ex, eh)
            ex, eh = part_errors(zs, a["z"][k + 1])
            assert ex < TOL and eh < TOL, ("zs", k, ex, eh)
            # free running
            z = O.sample_p_zs_given_zt(sd, cfg, s_arr, t_arr, z, nm, em, None, noise)
        ex, eh = part_errors(z, a["z"][4])
        assert ex < 1e-5 and eh < 1e-5, ("free", ex, eh)
        x, h = O.sample_p_xh_given_z0(sd, cfg, a["z"][4], nm, em, None, noise)
        ex, eh = part_errors(torch.cat([x, h], 2), a["xh0"])
        assert ex < TOL and eh < TOL


def test_full_sample_small_tamed_T1000():
    """End-to-end qm9/sampling.sample equivalent: identical global-generator draw order."""
    cfg, sd, a, _ = load_golden("small_tamed_sample_T1000")
    torch.manual_seed(int(a["torch_seed"][0]))
    with torch.no_grad():
        one_hot, charges, x, node_mask = O.sample_molecules(sd, cfg, a["nodes"].tolist(), 29)
    assert O.err_metric(x, a["x"]) < 1e-4
    assert torch.equal(one_hot.long(), a["one_hot"].long())
    assert torch.equal(charges.long(), a["charges"].long())


def test_nodes_distribution():
    cfg, sd, a, _ = load_golden("nodes_dist_qm9_seed0")
    from geoldm_b200.histograms import QM9_WITH_H_N_NODES
    torch.manual_seed(0)
    draws = O.nodes_distribution_sample(QM9_WITH_H_N_NODES, 64)
    assert torch.equal(draws, a["draws"])


def test_step_coefficients_match_tensor_path():
    cfg = O.QM9_CFG
    gamma = torch.from_numpy(O.noise_schedule_gamma(cfg))
    for s in (999, 500, 1, 0):
        a_ts, c_eps, c_noise = O.step_coefficients(gamma, 1000, s)
        assert torch.isfinite(a_ts) and torch.isfinite(c_eps) and torch.isfinite(c_noise)
        assert 0 < float(a_ts) <= 1.0


@pytest.mark.parametrize("name", ["v64_default", "v64_noatt", "v64_notanh", "v64_S2_noatt_notanh", "v64_mean",
                                  "v64_cond_latent2"])
def test_flag_variants_nf64(name):
    """Round-2 fixtures (oracle/make_golden_r2.py): every constructor flag at hidden_nf = 64, incl. 2- and 1-atom molecules."""
    cfg, sd, a, _ = load_golden(name)
    nm, em = O.build_masks(a["nodes"].tolist(), 29)
    ctx = a.get("context")
    with torch.no_grad():
        out = O.dynamics_forward(sd, cfg, a["t_vec"], a["z"], nm, em, ctx)
        dx, dh = O.decoder_forward(sd, cfg, a["z"], nm, em, ctx)
    ex, eh = part_errors(out, a["out"])
    assert ex < TOL and eh < TOL, (ex, eh)
    assert O.err_metric(dx, a["dec_x"]) < TOL and O.err_metric(dh, a["dec_h"]) < TOL


def test_full_size_trajectory_fixture_teacher_forced():
    """qm9_full_tamed_T1000 (complete 1000-step reference run at nf=256, 9 layers): the oracle reproduces eps_hat and z_s
    of stored steps early, in the middle and at the end of the trajectory, and p(x, h | z_0)."""
    cfg, sd, a, _ = load_golden("qm9_full_tamed_T1000")
    nodes = a["nodes"].tolist()
    bs, T, D = len(nodes), cfg.diffusion_steps, 3 + cfg.latent_nf
    nm, em = O.build_masks(nodes, 29)
    raw = torch.randn(T + 2, bs, 29, D, generator=torch.Generator().manual_seed(int(a["noise_seed"][0])))
    steps = a["steps"].tolist()
    with torch.no_grad():
        for k in (0, len(steps) // 2, len(steps) - 1):
            s = steps[k]
            s_arr = torch.full((bs, 1), float(s)) / T
            t_arr = torch.full((bs, 1), float(s + 1)) / T
            src = O.NoiseSource(raw.double()); src.k = T - s
            zs, eps = O.sample_p_zs_given_zt(sd, cfg, s_arr, t_arr, a["zt"][k], nm, em, None, src, True)
            ex, eh = part_errors(eps, a["eps"][k])
            assert ex < TOL and eh < TOL, ("eps", s, ex, eh)
            ex, eh = part_errors(zs, a["zs"][k])
            assert ex < TOL and eh < TOL, ("zs", s, ex, eh)
        src = O.NoiseSource(raw.double()); src.k = T + 1
        x, h = O.sample_p_xh_given_z0(sd, cfg, a["z0"], nm, em, None, src)
        ex, eh = part_errors(torch.cat([x, h], 2), a["xh0"])
        assert ex < TOL and eh < TOL, ("xh0", ex, eh)

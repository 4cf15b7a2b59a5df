"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference mount does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Every fixture stores the config (json), the weight seed (weights are re-created bit-identically by
oracle.geoldm_oracle.make_state_dict), the inputs and the reference outputs.  Nothing from the
reference's source is copied; it is only executed.
"""
from __future__ import annotations

import argparse
import contextlib
import dataclasses
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("GEOLDM_REFERENCE", "/root/reference")

from oracle import geoldm_oracle as O  # noqa: E402


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    sys.dont_write_bytecode = True
    # qm9/analyze.py and qm9/visualizer.py import matplotlib/imageio at module scope (absent here).
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import configs.datasets_config as dc
    import qm9.models as qm
    import qm9.sampling as qs
    return dc, qm, qs


def ref_args(cfg: O.OracleConfig, dataset: str):
    return argparse.Namespace(
        ae_path=None, cuda=False, include_charges=cfg.include_charges, context_node_nf=cfg.context_node_nf,
        conditioning=[], latent_nf=cfg.latent_nf, nf=cfg.nf, n_layers=cfg.n_layers, attention=cfg.attention,
        tanh=cfg.tanh, model="egnn_dynamics", norm_constant=cfg.norm_constant, inv_sublayers=cfg.inv_sublayers,
        sin_embedding=False, normalization_factor=cfg.normalization_factor,
        aggregation_method=cfg.aggregation_method, kl_weight=0.01, normalize_factors=list(cfg.normalize_factors),
        condition_time=cfg.condition_time, probabilistic_model="diffusion", diffusion_steps=cfg.diffusion_steps,
        diffusion_noise_schedule=cfg.diffusion_noise_schedule,
        diffusion_noise_precision=cfg.diffusion_noise_precision, diffusion_loss_type="l2", trainable_ae=False,
        ema_decay=0.999, dataset=dataset, remove_h=False)


def build_reference(cfg: O.OracleConfig, dataset: str, seed: int, tamed: bool, refmods, encoder=False,
                    trainable_ae=False):
    dc, qm, _ = refmods
    info = dc.get_dataset_info(dataset, False)
    args = ref_args(cfg, dataset)
    args.trainable_ae = trainable_ae
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(1234)
        model, _, _ = qm.get_latent_diffusion(args, "cpu", info, None)
    sd = O.make_state_dict(cfg, seed, tamed, encoder=encoder)
    ref_sd = model.state_dict()
    gam = ref_sd["gamma.gamma"].clone()
    missing = [k for k in sd if k not in ref_sd]
    assert not missing, missing
    for k, v in sd.items():
        assert ref_sd[k].shape == v.shape, (k, ref_sd[k].shape, v.shape)
        ref_sd[k] = v.clone()
    ref_sd["gamma.gamma"] = gam          # keep the reference's own table; compared separately
    model.load_state_dict(ref_sd)
    model.eval()
    return model, args, info, sd, gam


def random_latent(nodes, n_max, nf, gen, scale=1.0):
    nm, em = O.build_masks(nodes, n_max)
    z = torch.randn(len(nodes), n_max, 3 + nf, generator=gen) * nm
    zx = O.remove_mean_with_mask(z[:, :, :3], nm)
    return torch.cat([zx, z[:, :, 3:]], dim=2) * scale, nm, em


def save(name, cfg, dataset, seed, tamed, **arrays):
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    meta = dict(cfg=dataclasses.asdict(cfg), dataset=dataset, seed=seed, tamed=tamed,
                torch=torch.__version__, generator="oracle/make_golden.py")
    np.savez_compressed(out, meta=json.dumps(meta), **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                                       for k, v in arrays.items()})
    print("wrote", out, {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})


def main():
    refmods = import_reference()
    dc, qm, qs = refmods
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- G0: noise schedule table ---------------------------------------------------------------
    cfg = O.QM9_CFG
    model, args, info, sd, gam = build_reference(cfg, "qm9", 0, False, refmods)
    save("schedule_polynomial2_T1000", cfg, "qm9", 0, False, gamma=gam)

    # ---- G1: full-size QM9 denoiser forward (config 1/2 shapes, reduced batch) --------------------
    gen = torch.Generator().manual_seed(11)
    nodes = [5, 12, 18, 23, 29, 3, 19, 16]
    with torch.no_grad():
        arrays = {"nodes": np.array(nodes)}
        for tag, scale in (("s1", 1.0), ("s30", 30.0)):
            z, nm, em = random_latent(nodes, 29, cfg.latent_nf, gen, scale)
            t_scalar = torch.tensor([[0.5]])
            t_vec = torch.randint(0, 1001, (len(nodes), 1), generator=gen).float() / 1000.0
            arrays[f"z_{tag}"] = z
            arrays[f"t_vec_{tag}"] = t_vec
            arrays[f"out_tscalar_{tag}"] = model.dynamics._forward(t_scalar, z, nm, em, None)
            arrays[f"out_tvec_{tag}"] = model.dynamics._forward(t_vec, z, nm, em, None)
            arrays[f"out_t0_{tag}"] = model.dynamics._forward(torch.zeros(len(nodes), 1), z, nm, em, None)
        # decoder forward + decode on the same latent
        x_rec, h_rec = model.vae.decoder._forward(z / 30.0, nm, em, None)
        arrays["dec_in"] = z / 30.0
        arrays["dec_x"], arrays["dec_h"] = x_rec, h_rec
        x, h = model.vae.decode(z / 30.0, nm, em, None)
        arrays["decode_x"], arrays["decode_onehot"], arrays["decode_charges"] = x, h["categorical"], h["integer"]
    save("qm9_forward", cfg, "qm9", 0, False, **arrays)

    # ---- G2: teacher-forced first steps of the sampler at full size, injected noise ---------------
    nodes = [19, 23, 15, 27]
    bs, n_max, T = len(nodes), 29, cfg.diffusion_steps
    nm, em = O.build_masks(nodes, n_max)
    raw = torch.randn(6, bs, n_max, 3 + cfg.latent_nf, generator=torch.Generator().manual_seed(1234),
                      dtype=torch.float64)
    import equivariant_diffusion.utils as du
    k = {"i": 0}
    orig_x, orig_h = du.sample_center_gravity_zero_gaussian_with_mask, du.sample_gaussian_with_mask

    def inj_x(size, device, node_mask):
        r = raw[k["i"]][..., :3].float()
        return du.remove_mean_with_mask(r * node_mask, node_mask)

    def inj_h(size, device, node_mask):
        r = raw[k["i"]][..., 3:].float()
        k["i"] += 1
        return r * node_mask

    import equivariant_diffusion.en_diffusion as ed
    ed.utils.sample_center_gravity_zero_gaussian_with_mask = inj_x
    ed.utils.sample_gaussian_with_mask = inj_h
    try:
        with torch.no_grad():
            z = model.sample_combined_position_feature_noise(bs, n_max, nm)
            zs, epss = [z], []
            for s in reversed(range(T - 4, T)):
                s_arr = torch.full((bs, 1), s) / T
                t_arr = (torch.full((bs, 1), s) + 1) / T
                epss.append(model.phi(z, t_arr, nm, em, None))
                z = model.sample_p_zs_given_zt(s_arr, t_arr, z, nm, em, None)
                zs.append(z)
            k0 = k["i"]
            x0, h0 = model.sample_p_xh_given_z0(z, nm, em, None)
    finally:
        ed.utils.sample_center_gravity_zero_gaussian_with_mask = orig_x
        ed.utils.sample_gaussian_with_mask = orig_h
    save("qm9_sampler_steps", cfg, "qm9", 0, False, nodes=np.array(nodes), raw=raw, z=torch.stack(zs),
         eps=torch.stack(epss), xh0=torch.cat([x0, h0["integer"]], dim=2), raw_used=np.array([k0 + 1]))

    # ---- G3: small configs exercising every flag ---------------------------------------------------
    variants = {
        "small_default": O.OracleConfig(nf=32, n_layers=2),
        "small_S2_noatt_notanh": O.OracleConfig(nf=32, n_layers=2, inv_sublayers=2, attention=False, tanh=False,
                                                norm_constant=0.0, normalization_factor=100.0),
        "small_mean": O.OracleConfig(nf=64, n_layers=1, aggregation_method="mean"),
        "small_cond": O.OracleConfig(nf=64, n_layers=3, context_node_nf=1, include_charges=False,
                                     normalize_factors=(1.0, 8.0, 1.0)),
        "small_latent2": O.OracleConfig(nf=32, n_layers=2, latent_nf=2),
    }
    for name, c in variants.items():
        ds = "qm9_second_half" if c.context_node_nf else "qm9"
        m, _, _, _, _ = build_reference(c, ds, 3, False, refmods)
        gen = torch.Generator().manual_seed(5)
        nodes = [4, 9, 17, 29, 11]
        z, nm, em = random_latent(nodes, 29, c.latent_nf, gen)
        ctx = None
        if c.context_node_nf:
            ctx = torch.randn(len(nodes), 1, c.context_node_nf, generator=gen).repeat(1, 29, 1) * nm
        t_vec = torch.rand(len(nodes), 1, generator=gen)
        with torch.no_grad():
            out = m.dynamics._forward(t_vec, z, nm, em, ctx)
            dx, dh = m.vae.decoder._forward(z, nm, em, ctx)
        arrays = dict(nodes=np.array(nodes), z=z, t_vec=t_vec, out=out, dec_x=dx, dec_h=dh)
        if ctx is not None:
            arrays["context"] = ctx
        save(name, c, ds, 3, False, **arrays)

    # ---- G4: GEOM-Drugs shape (config 4), reduced batch -------------------------------------------
    c = O.GEOM_CFG
    m, _, _, _, _ = build_reference(c, "geom", 0, False, refmods)
    gen = torch.Generator().manual_seed(21)
    nodes = [44, 70, 181]
    z, nm, em = random_latent(nodes, 181, c.latent_nf, gen)
    with torch.no_grad():
        out = m.dynamics._forward(torch.tensor([[0.3]]), z, nm, em, None)
    save("geom_forward", c, "geom", 0, False, nodes=np.array(nodes), z=z, out=out)

    # ---- G5: complete 1000-step sample() through qm9/sampling.py on a small tamed model ------------
    c = O.OracleConfig(nf=64, n_layers=2)
    m, a, info_s, _, _ = build_reference(c, "qm9", 9, True, refmods)
    nodes = torch.tensor([7, 19, 29, 12])
    torch.manual_seed(77)
    with torch.no_grad():
        one_hot, charges, x, node_mask = qs.sample(a, "cpu", m, info_s, nodesxsample=nodes)
    save("small_tamed_sample_T1000", c, "qm9", 9, True, nodes=nodes, torch_seed=np.array([77]), one_hot=one_hot,
         charges=charges, x=x)

    # ---- G6: DistributionNodes draws ---------------------------------------------------------------
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        nd = qm.DistributionNodes(dc.qm9_with_h["n_nodes"])
    draws = nd.sample(64)
    save("nodes_dist_qm9_seed0", O.QM9_CFG, "qm9", 0, False, draws=draws)


if __name__ == "__main__":
    main()

"""Training-step fixtures (BASELINE config 5): loss and gradients of the UNMODIFIED reference on CPU.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_train.py

For each case the reference's ``qm9.losses.compute_loss_and_nll`` is run in train() mode (l2 objective, trainable
first stage) and its random draws (encoder noise, t, diffusion noise) are recorded so the CUDA path can be fed the same
ones; gradients of the batch-mean loss w.r.t. every dynamics / decoder parameter are stored (small cases: in full;
nf=192 L=9 conditional case: max / L2 / first-256-entries summaries).  The eval()-mode NLL estimate (two denoiser calls)
is stored for the small cases as well, and the first-stage (VAE-only) objective with its encoder gradients.
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
CHARGES = torch.tensor([1.0, 6.0, 7.0, 8.0, 9.0])


def make_batch(cfg, nodes, n_max, gen):
    nm, em = O.build_masks(nodes, n_max)
    bs = len(nodes)
    x = O.remove_mean_with_mask(torch.randn(bs, n_max, 3, generator=gen) * nm, nm)
    cat = torch.randint(0, cfg.n_atom_types, (bs, n_max), generator=gen)
    one_hot = torch.nn.functional.one_hot(cat, cfg.n_atom_types).float() * nm
    charges = (CHARGES[cat % 5].unsqueeze(2) * nm) if cfg.include_charges else torch.zeros(0)
    ctx = None
    if cfg.context_node_nf:
        ctx = torch.randn(bs, 1, cfg.context_node_nf, generator=gen).expand(-1, n_max, -1) * nm
    return x, one_hot, charges, nm, em, ctx


class Recorder:
    """Records the reference's random draws in call order."""

    def __init__(self, model):
        self.model, self.draws = model, {}
        self._orig = {}

    def __enter__(self):
        m = self.model
        self._orig = dict(vae=m.vae.sample_combined_position_feature_noise,
                          ld=m.sample_combined_position_feature_noise, randint=torch.randint)
        ld_names = iter(["eps_t", "eps_0"])

        def vae_noise(*a, **k):
            out = self._orig["vae"](*a, **k)
            self.draws["eps_enc"] = out.clone()
            return out

        def ld_noise(*a, **k):
            out = self._orig["ld"](*a, **k)
            self.draws[next(ld_names)] = out.clone()
            return out

        def randint(*a, **k):
            out = self._orig["randint"](*a, **k)
            self.draws["t_int"] = out.clone()
            return out

        m.vae.sample_combined_position_feature_noise = vae_noise
        m.sample_combined_position_feature_noise = ld_noise
        torch.randint = randint
        return self

    def __exit__(self, *exc):
        m = self.model
        m.vae.sample_combined_position_feature_noise = self._orig["vae"]
        m.sample_combined_position_feature_noise = self._orig["ld"]
        torch.randint = self._orig["randint"]


def summarise(g, full):
    names = sorted(g)
    if full:
        return {"g." + n: g[n] for n in names}
    return {"names": np.array(names),
            "gmax": np.array([g[n].abs().max().item() for n in names]),
            "gl2": np.array([g[n].double().norm().item() for n in names]),
            "ghead": np.stack([np.pad(g[n].flatten()[:HEAD].numpy(), (0, max(0, HEAD - g[n].numel()))) for n in names])}


def run_case(name, cfg, dataset, nodes, n_max, seed, refmods, full, with_eval, with_vae):
    dc, qm, qs = refmods
    import qm9.losses as ref_losses
    model, args, info, sd, gam = G.build_reference(cfg, dataset, seed, False, refmods, encoder=True, trainable_ae=True)
    nodes_dist = qm.DistributionNodes(info["n_nodes"])
    gen = torch.Generator().manual_seed(100 + seed)
    x, one_hot, charges, nm, em, ctx = make_batch(cfg, nodes, n_max, gen)
    h = {"categorical": one_hot, "integer": charges}
    arrays = {"nodes": np.array(nodes), "x": x, "one_hot": one_hot, "charges": charges}
    if ctx is not None:
        arrays["context"] = ctx
    # ---- train() mode: l2 objective, gradients ------------------------------------------------------
    model.train()
    torch.manual_seed(7)
    for p in model.parameters():
        p.grad = None
    with Recorder(model) as rec:
        nll, reg, _ = ref_losses.compute_loss_and_nll(args, model, nodes_dist, x, h, nm, em, ctx)
    nll.backward()
    arrays.update({"train_" + k: v for k, v in rec.draws.items()})
    arrays["train_loss"] = nll.detach()
    g = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    assert not any(n.startswith("vae.encoder") for n in g), "encoder is detached in EnLatentDiffusion.forward"
    arrays.update(summarise(g, full))
    if not full:
        # the same step in float64 (same weights, inputs and draws): the exact gradients both fp32 runs approximate
        m64 = model.double()
        for p in m64.parameters():
            p.grad = None
        d64 = {k: v.double() for k, v in rec.draws.items()}
        seq = iter([d64["eps_t"]])
        o_v, o_l, o_r = m64.vae.sample_combined_position_feature_noise, m64.sample_combined_position_feature_noise, torch.randint
        m64.vae.sample_combined_position_feature_noise = lambda *a, **k: d64["eps_enc"]
        m64.sample_combined_position_feature_noise = lambda *a, **k: next(seq)
        torch.randint = lambda *a, **k: rec.draws["t_int"]
        try:
            h64 = {k: v.double() for k, v in h.items()}
            nll64, _, _ = ref_losses.compute_loss_and_nll(args, m64, nodes_dist, x.double(), h64, nm.double(),
                                                          em.double(), None if ctx is None else ctx.double())
            nll64.backward()
        finally:
            m64.vae.sample_combined_position_feature_noise, m64.sample_combined_position_feature_noise = o_v, o_l
            torch.randint = o_r
        g64 = {n: p.grad.detach().clone() for n, p in m64.named_parameters() if p.grad is not None}
        s64 = summarise(g64, False)
        arrays.update({"gmax64": s64["gmax"], "gl264": s64["gl2"], "ghead64": s64["ghead"],
                       "train_loss64": nll64.detach()})
        model.float()
    with Recorder(model) as rec, torch.no_grad():
        torch.manual_seed(8)
        per_mol = model(x, h, nm, em.view(len(nodes), -1), ctx)
    arrays.update({"train2_" + k: v for k, v in rec.draws.items()})
    arrays["train2_per_mol"] = per_mol
    # ---- eval() mode: NLL estimate (t0_always) -----------------------------------------------------------
    if with_eval:
        model.eval()
        with Recorder(model) as rec, torch.no_grad():
            torch.manual_seed(9)
            per_mol = model(x, h, nm, em.view(len(nodes), -1), ctx)
        arrays.update({"eval_" + k: v for k, v in rec.draws.items()})
        arrays["eval_per_mol"] = per_mol
    # ---- first stage alone (EnHierarchicalVAE.forward) ------------------------------------------------------
    if with_vae:
        vae = model.vae
        vae.train()
        for p in vae.parameters():
            p.grad = None
        with Recorder(model) as rec:
            torch.manual_seed(10)
            loss = vae(x, h, nm, em, ctx)
        loss.mean().backward()
        arrays["vae_eps_enc"] = rec.draws["eps_enc"]
        arrays["vae_per_mol"] = loss.detach()
        arrays.update({"vg." + n: p.grad.detach().clone() for n, p in vae.named_parameters() if p.grad is not None})
    G.save(name, cfg, dataset, seed, False, **arrays)


def main():
    refmods = G.import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    run_case("train_small", O.OracleConfig(nf=32, n_layers=2), "qm9", [5, 9, 7, 3, 8, 4], 9, 4, refmods,
             full=True, with_eval=True, with_vae=True)
    run_case("train_small_cond", O.OracleConfig(nf=32, n_layers=2, context_node_nf=1, include_charges=False,
                                                normalize_factors=(1.0, 8.0, 1.0)),
             "qm9_second_half", [6, 4, 9, 9], 9, 5, refmods, full=True, with_eval=True, with_vae=False)
    run_case("train_qm9cond", O.QM9_COND_CFG, "qm9_second_half", [17, 23, 9, 29, 19, 12, 21, 4], 29, 0, refmods,
             full=False, with_eval=False, with_vae=False)


if __name__ == "__main__":
    main()

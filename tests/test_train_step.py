"""BASELINE config 5: the training objective (EnLatentDiffusion.forward through qm9/losses.py:compute_loss_and_nll) and
its gradients against the unmodified reference with the reference's own random draws injected
(tests/golden/train_*.npz from oracle/make_golden_train.py).  Metric per tensor: max|a-b| / max|b|, tolerance 1e-5
for losses, GRAD_TOL for gradients (stated below)."""
import argparse

import numpy as np
import pytest
import torch

from oracle import geoldm_oracle as O
from tests.helpers import build_cuda_model, load_golden

LOSS_TOL = 1e-5
GRAD_TOL = 3e-5        # measured 1e-6 (small cases) .. 1.3e-5 (nf=192, 9 blocks, after subtracting the reference's own fp32 noise)
# vs the float64 run of the reference (full-size case), two classes of tensors:
#  * more than one element: 1.5e-5 (measured over 30 runs on B200: 2e-6 .. 7e-6; the rest of the margin is for the order of
#    the fp32 atomics in the scatter kernels, which differs from run to run);
#  * single-element tensors - the attention biases, each a CANCELLING sum of ~20 k signed per-edge terms that leaves a
#    5e-6-sized value: 5e-5.  Their own reduction is deterministic and in double (train.cu, dbw_scratch), but their inputs
#    carry the 1e-7 order-of-atomics noise of the upstream scatter adds, which the cancellation amplifies: measured 3e-6 ..
#    2.3e-5 over 30 runs (3.1e-5 before the deterministic reduction).  The reference's own fp32 run is 3.6e-5 away from its
#    fp64 run on one of them.
GRAD_TOL64 = 1.5e-5
GRAD_TOL64_SCALAR = 5e-5


def _rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def _inputs(A, cfg, device):
    nodes = A["nodes"].tolist()
    nm, em = O.build_masks(nodes, A["x"].shape[1])
    h = {"categorical": A["one_hot"].to(device), "integer": A["charges"].to(device)}
    ctx = A["context"].to(device) if "context" in A else None
    return A["x"].to(device), h, nm.to(device), em.to(device), ctx


def _draws(A, prefix, device):
    return {k[len(prefix):]: v.to(device) for k, v in A.items() if k.startswith(prefix) and k[len(prefix):] in
            ("eps_enc", "t_int", "eps_t", "eps_0")}


def _nodes_dist(meta):
    from geoldm_b200.histograms import HISTOGRAMS
    from geoldm_b200.models import DistributionNodes
    return DistributionNodes(HISTOGRAMS[meta["dataset"]])


def _force_autograd_everywhere(model, monkeypatch):
    """CPU-only check of losses.py + the autograd graph: route every wrapper through train.py's library-op graph.
    (The product refuses CPU tensors; this patches the test's own model instance.)"""
    from geoldm_b200 import dynamics
    from geoldm_b200 import train as _train
    monkeypatch.setattr(_train, "_CPU_GRAPH_CHECK", True)        # test seam, see train.allow_cpu_graph_check
    monkeypatch.setattr(dynamics._EgnnWrapper, "_check_inputs", lambda self, xh, nm: None)
    monkeypatch.setattr(dynamics._EgnnWrapper, "_wants_grad", lambda self, xh, ctx: True)


def _check_case(name, device, monkeypatch=None):
    from geoldm_b200.losses import compute_loss_and_nll
    cfg, sd, A, meta = load_golden(name, encoder=True)
    model = build_cuda_model(cfg, sd, device=device, trainable_ae=True)
    if monkeypatch is not None:
        _force_autograd_everywhere(model, monkeypatch)
    x, h, nm, em, ctx = _inputs(A, cfg, device)
    args = argparse.Namespace(probabilistic_model="diffusion")
    model.train()
    assert model.vae.training
    loss, _, _ = compute_loss_and_nll(args, model, _nodes_dist(meta), x, h, nm, em, ctx,
                                      draws=_draws(A, "train_", device))
    e_loss = _rel(loss, A["train_loss"])
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert not any(n.startswith("vae.encoder") for n in grads)
    worst, wname = 0.0, None
    if "names" in A:
        names = [str(s) for s in np.asarray(A["names"])]
        assert sorted(grads) == names
        # Same well-posed form as the forward gate (helpers.assert_parity): the reference's own fp32 gradients carry
        # rounding noise (up to 3.6e-5 on a 5e-6-sized attention-bias gradient that is a cancelling sum over all
        # edges), so each tensor is gated by  err(ours, ref32) <= GRAD_TOL + err(ref32, ref64)  and
        # err(ours, ref64) <= GRAD_TOL64.
        worst64, floor_at_worst, wname64 = 0.0, 0.0, None
        worst64_scalar = 0.0
        for i, n in enumerate(names):
            gp = grads[n]
            head = gp.flatten()[:256].double().cpu()

            def err(mx, l2, hd):
                mx, l2 = float(mx), float(l2)
                return max(abs(float(gp.abs().max()) - mx) / mx, abs(float(gp.double().norm()) - l2) / l2,
                           float((head - hd[:head.numel()].double()).abs().max()) / mx)
            e32 = err(A["gmax"][i], A["gl2"][i], A["ghead"][i])
            e64 = err(A["gmax64"][i], A["gl264"][i], A["ghead64"][i])
            floor = max(abs(float(A["gmax"][i]) - float(A["gmax64"][i])) / float(A["gmax64"][i]),
                        float((A["ghead"][i].double() - A["ghead64"][i].double()).abs().max()) / float(A["gmax64"][i]))
            if gp.numel() == 1:
                worst64_scalar = max(worst64_scalar, e64)
            elif e64 > worst64:
                worst64, wname64 = e64, n
            if e32 - floor > worst - floor_at_worst:
                worst, wname, floor_at_worst = e32, n, floor
        print(f"[train] {name}: worst grad vs fp64 reference {worst64:.2e} ({wname64}), single-element tensors "
              f"{worst64_scalar:.2e}; vs fp32 reference {worst:.2e} "
              f"(reference's own fp32 noise there {floor_at_worst:.2e})")
        assert worst64 < GRAD_TOL64 and worst64_scalar < GRAD_TOL64_SCALAR
        worst = max(0.0, worst - floor_at_worst)
    else:
        ref = {k[2:]: v for k, v in A.items() if k.startswith("g.")}
        assert sorted(grads) == sorted(ref)
        for n in ref:
            e = _rel(grads[n], ref[n])
            if e > worst:
                worst, wname = e, n
    with torch.no_grad():
        per_mol = model(x, h, nm, em.view(len(A["nodes"]), -1), ctx, draws=_draws(A, "train2_", device))
    e_pm = _rel(per_mol, A["train2_per_mol"])
    msg = f"[train] {name} ({device}): loss {e_loss:.2e}, per-molecule {e_pm:.2e}, worst grad {wname} = {worst:.2e}"
    e_ev = 0.0
    if "eval_per_mol" in A:
        model.eval()
        with torch.no_grad():
            ev = model(x, h, nm, em.view(len(A["nodes"]), -1), ctx, draws=_draws(A, "eval_", device))
        e_ev = _rel(ev, A["eval_per_mol"])
        msg += f", eval NLL {e_ev:.2e}"
    print(msg)
    assert e_loss < LOSS_TOL and e_pm < LOSS_TOL and e_ev < LOSS_TOL and worst < GRAD_TOL
    if any(k.startswith("vg.") for k in A):
        vae = model.vae.train()
        for p in vae.parameters():
            p.grad = None
        lv = vae(x, h, nm, em, ctx, draws={"eps_enc": A["vae_eps_enc"].to(device)})
        e_v = _rel(lv, A["vae_per_mol"])
        lv.mean().backward()
        ref = {k[3:]: v for k, v in A.items() if k.startswith("vg.")}
        got = {n: p.grad for n, p in vae.named_parameters() if p.grad is not None}
        assert sorted(got) == sorted(ref)
        wv = max(_rel(got[n], ref[n]) for n in ref)
        print(f"[train] {name} first stage: loss {e_v:.2e}, worst grad {wv:.2e}")
        assert e_v < LOSS_TOL and wv < GRAD_TOL


@pytest.mark.parametrize("name", ["train_small", "train_small_cond"])
def test_training_objective_math_cpu(name, monkeypatch):
    _check_case(name, "cpu", monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["train_small", "train_small_cond", "train_qm9cond"])
def test_training_step_cuda(name):
    _check_case(name, "cuda")


@pytest.mark.gpu
def test_optimisation_steps_reduce_the_loss_and_refresh_the_inference_kernels():
    """A few train_step calls on a fixed batch: the l2 objective goes down, and the fused inference kernels see the
    updated weights afterwards (their packed images are keyed on parameter versions)."""
    import copy
    from geoldm_b200 import training
    cfg, sd, A, meta = load_golden("train_small", encoder=True)
    model = build_cuda_model(cfg, sd, device="cuda", mma_mode="fp32", trainable_ae=True)
    x, h, nm, em, ctx = _inputs(A, cfg, "cuda")
    draws = _draws(A, "train_", "cuda")
    args = argparse.Namespace(probabilistic_model="diffusion", lr=2e-3, clip_grad=True, ema_decay=0.9, ode_regularization=0.0)
    z = torch.randn(len(A["nodes"]), x.shape[1], 4, device="cuda") * nm
    t = torch.full((len(A["nodes"]), 1), 0.5, device="cuda")
    model.eval()
    with torch.no_grad():
        before = model.dynamics._forward(t, z, nm, em, None).clone()
    model_ema = copy.deepcopy(model)
    optim = training.get_optim(args, model)
    queue = training.Queue()
    queue.add(3000.0)
    losses = []
    for _ in range(6):
        nll, gn = training.train_step(args, model, optim, _nodes_dist(meta), x, h, nm, em, ctx, gradnorm_queue=queue,
                                      model_ema=model_ema, ema=training.EMA(args.ema_decay), draws=draws)
        losses.append(float(nll))
    print("[train] losses over 6 steps on a fixed batch:", [round(v, 4) for v in losses])
    assert losses[-1] < losses[0]
    model.eval()
    with torch.no_grad():
        after = model.dynamics._forward(t, z, nm, em, None)
        ema_out = model_ema.eval().dynamics._forward(t, z, nm, em, None)
    assert float((after - before).abs().max()) > 1e-4          # the inference path uses the new weights
    assert float((ema_out - before).abs().max()) > 0 and float((ema_out - after).abs().max()) > 0


def test_device_grad_clip_matches_host_queue():
    """DeviceGradClip (ring buffer + threshold + clip on device tensors) == utils.py's Queue + gradient_clipping
    (training.Queue / training.gradient_clipping) step by step, including the 50-entry window and the clipped branch."""
    from geoldm_b200 import training
    torch.manual_seed(3)
    lin_a, lin_b = torch.nn.Linear(7, 5), torch.nn.Linear(7, 5)
    lin_b.load_state_dict(lin_a.state_dict())
    q = training.Queue()
    q.add(3000.0)
    dc = training.DeviceGradClip(torch.device("cpu"))
    for it in range(70):
        scale = 50.0 if it % 9 == 4 else 1.0 + 0.1 * (it % 5)          # occasional spikes: the clipped branch
        ga = [torch.randn_like(p) * scale for p in lin_a.parameters()]
        for p, g in zip(lin_a.parameters(), ga):
            p.grad = g.clone()
        gb = [g.clone() for g in ga]
        n_host = training.gradient_clipping(lin_a, q)
        n_dev = dc.clip_(gb)
        assert abs(float(n_host) - float(n_dev)) <= 1e-5 * float(n_host)
        for p, g in zip(lin_a.parameters(), gb):
            assert torch.allclose(p.grad, g, rtol=1e-5, atol=1e-7)
        assert int(dc.count) == len(q)
        assert abs(float(dc.threshold()) - (1.5 * q.mean() + 2 * q.std())) <= 1e-4 * (1.5 * q.mean() + 2 * q.std())


@pytest.mark.gpu
def test_graphed_train_step_equals_eager():
    """training.GraphedTrainStep (whole step captured as one CUDA graph, device-side clipping, capturable AdamW) follows the
    eager train_step: same NLL trajectory from the same seed, same weights after 1 eager + 5 replayed steps."""
    import copy
    from geoldm_b200 import training
    cfg, sd, A, meta = load_golden("train_small", encoder=True)
    args = argparse.Namespace(probabilistic_model="diffusion", lr=1e-3, clip_grad=True, ema_decay=0.9, ode_regularization=0.0)

    def run(graphed):
        model = build_cuda_model(cfg, sd, device="cuda", mma_mode="fp32", trainable_ae=True)
        x, h, nm, em, ctx = _inputs(A, cfg, "cuda")
        model_ema = copy.deepcopy(model)
        ema = training.EMA(args.ema_decay)
        nd = _nodes_dist(meta)
        torch.manual_seed(17)
        out = []
        if graphed:
            optim = training.get_optim(args, model, capturable=True)
            g = training.GraphedTrainStep(args, model, optim, nd, x, h, nm, em, ctx, model_ema=model_ema, ema=ema, warmup=1)    # = one eager step
            for _ in range(5):
                out.append(float(g(x, h, ctx)[0]))
        else:
            optim = training.get_optim(args, model)
            q = training.Queue()
            q.add(3000.0)
            buckets = training.FlatGradBuckets(model)
            for _ in range(6):
                out.append(float(training.train_step(args, model, optim, nd, x, h, nm, em, ctx, gradnorm_queue=q,
                                                     model_ema=model_ema, ema=ema, buckets=buckets)[0]))
            out = out[1:]
        return out, [p.detach().clone() for p in model.parameters()], [p.detach().clone() for p in model_ema.parameters()]

    l_e, p_e, m_e = run(False)
    l_g, p_g, m_g = run(True)
    print("[graph] eager", [round(v, 5) for v in l_e], "graphed", [round(v, 5) for v in l_g])
    # step 0 draws from the same generator state; later steps differ only through fp32 atomics in the backward kernels
    assert abs(l_e[0] - l_g[0]) <= 1e-5 * abs(l_e[0])
    for a, b in zip(l_e, l_g):
        assert abs(a - b) <= 1e-4 * abs(a)              # measured: identical to 7 digits
    worst = max(float((a - b).abs().max() / (a.abs().max() + 1e-12)) for a, b in zip(p_e, p_g))
    worst_ema = max(float((a - b).abs().max() / (a.abs().max() + 1e-12)) for a, b in zip(m_e, m_g))
    print(f"[graph] weights after 5 steps: eager vs graphed rel {worst:.2e}, EMA {worst_ema:.2e}")
    assert worst < 5e-4 and worst_ema < 5e-4             # measured 2.1e-5 / 2.9e-5 (fp32 atomics in the backward kernels)


@pytest.mark.gpu
def test_fused_adamw_ema_matches_library_optimiser():
    """training.FusedAdamWEMA (geoldm_adamw_ema_step: AdamW(amsgrad) + EMA as one multi-tensor launch on the library
    optimiser's own state tensors) against torch.optim.AdamW + EMA.update_model_average on the same gradients, 6 updates,
    tensors from 1 to 70 000 elements (several chunks); `optim.state_dict()` stays usable and the step counters advance."""
    import copy
    from geoldm_b200 import training
    args = argparse.Namespace(lr=2e-3)
    torch.manual_seed(3)

    def make():
        torch.manual_seed(5)
        net = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, device="cuda")) for s in ((1,), (17, 5), (70000,), (4096,), (64, 65))])
        return net, copy.deepcopy(net)

    grads = [[torch.randn_like(p) * (0.1 + k) for p in make()[0]] for k in range(7)]
    ema = training.EMA(0.9)
    # library path
    net_a, ema_a = make()
    opt_a = training.get_optim(args, net_a, capturable=True)
    for k in range(7):
        for p, g in zip(net_a, grads[k]):
            p.grad = g.clone()
        opt_a.step()
        ema.update_model_average(ema_a, net_a)
    # fused path: first update through the library (creates the state), then six fused updates
    net_b, ema_b = make()
    opt_b = training.get_optim(args, net_b, capturable=True)
    for p, g in zip(net_b, grads[0]):
        p.grad = g.clone()
    opt_b.step()
    ema.update_model_average(ema_b, net_b)
    fused = training.FusedAdamWEMA(opt_b, net_b, ema_b, ema)
    for k in range(1, 7):
        for p, g in zip(net_b, grads[k]):
            p.grad.copy_(g)
        fused.check_attached(opt_b)
        fused.step()
    rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-12))
    w = max(rel(a, b) for a, b in zip(net_b, net_a))
    e = max(rel(a, b) for a, b in zip(ema_b, ema_a))
    sa, sb = opt_a.state_dict()["state"], opt_b.state_dict()["state"]
    st = max(rel(sb[i][k], sa[i][k]) for i in sa for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"))
    steps = {float(sb[i]["step"]) for i in sb}
    print(f"[fused optim] weights {w:.2e}, EMA {e:.2e}, moments {st:.2e}, steps {steps}")
    assert w < 2e-6 and e < 2e-6 and st < 2e-6 and steps == {7.0}

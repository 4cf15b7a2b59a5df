"""Config-5 check on real GPUs: one conditional QM9 training step (nf=192, 9 blocks, batch 64 per GPU) per rank with the
NCCL gradient all-reduce; verifies that all ranks hold identical weights after the step and that the all-reduced
gradient equals the single-process gradient of the concatenated batch; prints step timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/train_check.py [--steps 5]
"""
import argparse
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from geoldm_b200.histograms import HISTOGRAMS                      # noqa: E402
from geoldm_b200.models import DistributionNodes, get_latent_diffusion   # noqa: E402
from geoldm_b200 import losses, training                           # noqa: E402
from geoldm_b200.sampling import build_masks                       # noqa: E402


def make_args(nf, n_layers):
    return argparse.Namespace(
        ae_path=None, cuda=True, include_charges=False, context_node_nf=1, conditioning=[], latent_nf=1, nf=nf,
        n_layers=n_layers, attention=True, tanh=True, model="egnn_dynamics", norm_constant=1, inv_sublayers=1,
        sin_embedding=False, normalization_factor=1, aggregation_method="sum", kl_weight=0.01,
        normalize_factors=[1, 8, 1], condition_time=True, probabilistic_model="diffusion", diffusion_steps=1000,
        diffusion_noise_schedule="polynomial_2", diffusion_noise_precision=1e-5, diffusion_loss_type="l2",
        trainable_ae=True, ema_decay=0.999, dataset="qm9_second_half", remove_h=False, mma_mode="3xf16", lr=1e-4,
        clip_grad=True, ode_regularization=0.0)


def synth_batch(nodes, n_max, device, gen):
    nodes_t = torch.as_tensor(nodes)
    nm, em = build_masks(nodes_t, n_max, device)
    bs = len(nodes)
    x = torch.randn(bs, n_max, 3, generator=gen).to(device) * nm
    x = losses.remove_mean_with_mask(x, nm)
    cat = torch.randint(0, 5, (bs, n_max), generator=gen).to(device)
    one_hot = torch.nn.functional.one_hot(cat, 5).float() * nm
    ctx = torch.randn(bs, 1, 1, generator=gen).to(device).expand(-1, n_max, -1) * nm
    draws = {"eps_enc": losses.masked_noise(bs, n_max, 3, 1, nm), "eps_t": losses.masked_noise(bs, n_max, 3, 1, nm),
             "t_int": torch.randint(0, 1001, (bs, 1), generator=gen).to(device)}
    return x, {"categorical": one_hot, "integer": torch.zeros(0, device=device)}, nm, em, ctx, draws


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--nf", type=int, default=192)
    ap.add_argument("--n-layers", type=int, default=9)
    ap.add_argument("--graph", action="store_true", help="time training.GraphedTrainStep instead of the eager train_step")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args = make_args(a.nf, a.n_layers)
    info = {"atom_decoder": list(range(5)), "n_nodes": HISTOGRAMS["qm9_second_half"], "max_n_nodes": 29}
    torch.manual_seed(0)                                    # identical init on every rank
    model, nodes_dist, _ = get_latent_diffusion(args, dev, info, None)
    rng = np.random.default_rng(5)
    sizes = np.array(list(HISTOGRAMS["qm9_second_half"].keys()))
    prob = np.array(list(HISTOGRAMS["qm9_second_half"].values()), dtype=np.float64)
    all_nodes = rng.choice(sizes, size=world * a.batch, p=prob / prob.sum()).tolist()
    gen = torch.Generator().manual_seed(11)
    full = synth_batch(all_nodes, 29, dev, gen)              # same on every rank (CPU generator + seeded CUDA draws)
    torch.manual_seed(1)
    full[5]["eps_enc"] = losses.masked_noise(len(all_nodes), 29, 3, 1, full[2])
    full[5]["eps_t"] = losses.masked_noise(len(all_nodes), 29, 3, 1, full[2])
    sl = slice(rank * a.batch, (rank + 1) * a.batch)

    def shard(sel):
        x, h, nm, em, ctx, draws = full
        bs = nm[sel].shape[0]
        return (x[sel], {"categorical": h["categorical"][sel], "integer": h["integer"]}, nm[sel],
                em.view(len(all_nodes), -1)[sel].reshape(-1, 1), ctx[sel], {k: v[sel] for k, v in draws.items()})

    report = {"world": world, "batch_per_gpu": a.batch, "nf": a.nf, "n_layers": a.n_layers}
    # ---- gradient check: all-reduced shard gradients == gradients of the whole batch ------------------------------
    model.train()
    x, h, nm, em, ctx, draws = shard(sl)
    nll, _, _ = losses.compute_loss_and_nll(args, model, nodes_dist, x, h, nm, em, ctx, draws=draws)
    nll.backward()
    buckets = training.gradient_buckets(model)
    nbytes = training.allreduce_gradients(buckets) if world > 1 else 0
    mine = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    x, h, nm, em, ctx, draws = shard(slice(0, len(all_nodes)))
    nll_all, _, _ = losses.compute_loss_and_nll(args, model, nodes_dist, x, h, nm, em, ctx, draws=draws)
    nll_all.backward()
    errs = {n: float((mine[n] - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-30))
            for n, p in model.named_parameters() if p.grad is not None}
    worst_name = max(errs, key=errs.get)
    worst = errs[worst_name]
    report.update(allreduce_bytes=nbytes, grad_vs_whole_batch=worst, worst_tensor=worst_name)
    if not worst < 2e-5:
        p = dict(model.named_parameters())[worst_name]
        raise AssertionError(f"{worst_name}: rel {worst:.3e}, |whole-batch grad| max {float(p.grad.abs().max()):.3e}, "
                             f"|all-reduced shard grad| max {float(mine[worst_name].abs().max()):.3e}")
    model.zero_grad(set_to_none=True)
    del nll_all, mine, nll                         # no autograd graph of the default stream survives into a capture
    # ---- timed steps --------------------------------------------------------------------------------------------------
    graphed = "--graph" in sys.argv                # the step as ONE captured CUDA graph (training.GraphedTrainStep)
    optim = training.get_optim(args, model, capturable=graphed)
    buckets = training.FlatGradBuckets(model)      # timed steps: flat gradients, all-reduce overlapped with backward
    q = training.Queue()
    q.add(3000.0)
    model_ema = copy.deepcopy(model)
    ema = training.EMA(args.ema_decay)
    x, h, nm, em, ctx, draws = shard(sl)
    gstep = None
    if graphed:
        gstep = training.GraphedTrainStep(args, model, optim, nodes_dist, x, h, nm, em, ctx, model_ema=model_ema, ema=ema,
                                          buckets=buckets)
    report["graphed"] = graphed
    times = []
    for it in range(a.steps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if gstep is not None:
            nll, gn = gstep(x, h, ctx)
        else:
            nll, gn = training.train_step(args, model, optim, nodes_dist, x, h, nm, em, ctx, gradnorm_queue=q,
                                          model_ema=model_ema, ema=ema, buckets=buckets)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = torch.tensor([float(np.mean(times))], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # ---- ranks in lock-step ---------------------------------------------------------------------------------------------
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    if world > 1:
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        assert all(torch.equal(g, gathered[0]) for g in gathered), gathered
    report.update(ms_per_step=float(ms), molecules_per_s=world * a.batch / float(ms) * 1e3, last_nll=float(nll),
                  lockstep=True)
    if rank == 0:
        print(json.dumps(report))
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/train_check_n{world}.json", "w") as f:
            json.dump(report, f)
    if gstep is not None:
        gstep.close()                               # the graph holds NCCL work: release it before the process group
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

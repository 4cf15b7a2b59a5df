"""Kernel count and summed GPU time of one config-5 training step (torch.profiler), vs. its wall time."""
import os, sys, copy, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from geoldm_b200 import losses, training
from geoldm_b200.histograms import HISTOGRAMS
from geoldm_b200.models import get_latent_diffusion
from geoldm_b200.sampling import build_masks
dev = torch.device("cuda:0")
bs = 64
args = bench.qm9_args("3xf16")
args.include_charges, args.context_node_nf, args.nf, args.normalize_factors = False, 1, 192, [1, 8, 1]
args.trainable_ae, args.dataset, args.lr, args.clip_grad, args.ode_regularization = True, "qm9_second_half", 1e-4, True, 0.0
hist = HISTOGRAMS["qm9_second_half"]
info = {"atom_decoder": list(range(5)), "n_nodes": hist, "max_n_nodes": 29}
torch.manual_seed(0)
model, nodes_dist, _ = get_latent_diffusion(args, dev, info, None)
nodes = bench.histogram_nodes(hist, bs, seed=5)
gen = torch.Generator().manual_seed(11)
nm, em = build_masks(torch.as_tensor(nodes), 29, dev)
x = losses.remove_mean_with_mask(torch.randn(bs, 29, 3, generator=gen).to(dev) * nm, nm)
one_hot = torch.nn.functional.one_hot(torch.randint(0, 5, (bs, 29), generator=gen).to(dev), 5).float() * nm
ctx = torch.randn(bs, 1, 1, generator=gen).to(dev).expand(-1, 29, -1) * nm
h = {"categorical": one_hot, "integer": torch.zeros(0, device=dev)}
optim = training.get_optim(args, model)
q = training.Queue(); q.add(3000.0)
model_ema = copy.deepcopy(model); ema = training.EMA(args.ema_decay)
buckets = training.FlatGradBuckets(model)
def step():
    return training.train_step(args, model, optim, nodes_dist, x, h, nm, em, ctx, gradnorm_queue=q, model_ema=model_ema, ema=ema, buckets=buckets)
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev)
import collections
by = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    by[e.name[:60]][0] += 1; by[e.name[:60]][1] += e.device_time
print(json.dumps({"kernels": len(ev), "gpu_busy_ms": tot / 1e3}))
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[0]:5d} {v[1]/1e3:8.3f} ms  {k}")

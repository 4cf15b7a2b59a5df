"""Config-5 training step: eager train_step vs the captured GraphedTrainStep (ms per step, NLL trajectory)."""
import os, sys, copy, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from geoldm_b200 import losses, training
from geoldm_b200.histograms import HISTOGRAMS
from geoldm_b200.models import get_latent_diffusion
from geoldm_b200.sampling import build_masks
dev = torch.device("cuda:0")
bs = 64
def setup():
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
    return args, model, nodes_dist, x, h, nm, em, ctx
def timed(fn, n):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
res = {}
# eager
args, model, nodes_dist, x, h, nm, em, ctx = setup()
optim = training.get_optim(args, model); q = training.Queue(); q.add(3000.0)
model_ema = copy.deepcopy(model); ema = training.EMA(args.ema_decay); buckets = training.FlatGradBuckets(model)
step = lambda: training.train_step(args, model, optim, nodes_dist, x, h, nm, em, ctx, gradnorm_queue=q, model_ema=model_ema, ema=ema, buckets=buckets)[0]
for _ in range(3): step()
ms, out = timed(step, 20)
res["eager_ms"] = ms; res["eager_nll"] = [round(float(v), 3) for v in out[::4]]
# graphed
args, model, nodes_dist, x, h, nm, em, ctx = setup()
optim = training.get_optim(args, model, capturable=True)
model_ema = copy.deepcopy(model); ema = training.EMA(args.ema_decay)
g = training.GraphedTrainStep(args, model, optim, nodes_dist, x, h, nm, em, ctx, model_ema=model_ema, ema=ema)
p0 = next(model.dynamics.parameters()).detach().clone()
ms, out = timed(lambda: g(x)[0].clone(), 20)
res["graph_ms"] = ms; res["graph_nll"] = [round(float(v), 3) for v in out[::4]]
res["param_moved"] = float((next(model.dynamics.parameters()) - p0).abs().max())
res["grad_norm_hist"] = [round(float(v), 2) for v in g.clip.hist[:6]]
print(json.dumps(res))

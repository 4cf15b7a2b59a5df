"""Per-kernel GPU time of ONE replay of the captured config-5 training step (torch.profiler over graph.replay())."""
import os, sys, copy, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_profile.py")).read().split("optim = training.get_optim")[0]
exec(compile(src, "head", "exec"))
optim = training.get_optim(args, model, capturable=True)
model_ema = copy.deepcopy(model); ema = training.EMA(args.ema_decay)
g = training.GraphedTrainStep(args, model, optim, nodes_dist, x, h, nm, em, ctx, model_ema=model_ema, ema=ema)
for _ in range(3): g(x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(x); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
by = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    by[e.name[:70]][0] += 1; by[e.name[:70]][1] += e.device_time
print(json.dumps({"kernels": len(ev), "gpu_busy_ms": sum(e.device_time for e in ev) / 1e3}))
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v[0]:5d} {v[1]/1e3:8.3f} ms  {k}")

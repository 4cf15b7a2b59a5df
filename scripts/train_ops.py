"""aten-level op counts of one eager config-5 training step (which library ops make up the ~1900 element-wise launches)."""
import os, sys, runpy
sys.argv = ["train_profile.py"]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_profile.py")).read()
src = src.split("from torch.profiler import profile")[0]
exec(compile(src, "train_profile_head", "exec"))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
rows = [(e.key, e.count, e.self_device_time_total / 1e3) for e in prof.key_averages() if e.key.startswith(("aten::", "autograd::", "Optimizer")) and e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[2])
for k, c, t in rows[:28]:
    print(f"{c:5d} {t:8.3f} ms  {k}")

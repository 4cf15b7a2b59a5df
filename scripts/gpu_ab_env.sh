#!/bin/bash
# A/B of environment switches on the in-tree library: scripts/gpu_ab_env.sh "GEOLDM_TC_STAGE=0" "GEOLDM_TC_STAGE=1" ...
mkdir -p gpurun_out
for envs in "$@"; do
  env $envs python bench.py --steps 100 --warmup 20 --no-e2e --no-cpu-baseline --no-extra 2>gpurun_out/abq.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$envs', round(d['value'],1), 'mol/s', round(d['ms_per_step'],4), 'ms/step  gcl', round(d['roofline']['kernel_ms'],4), d['clocks']['sm_mhz'])" || tail -5 gpurun_out/abq.err
done

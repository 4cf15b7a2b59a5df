#!/bin/bash
# One gpurun call: GPU parity tests, then A/B of library variants (GEOLDM_B200_LIB) on the bench step and the cycle counters.
# usage: scripts/gpu_ab.sh [variant.so ...]   (the in-tree library always runs first)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/ab_pytest_full.log 2>&1
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/ab_pytest_full.log > gpurun_out/ab_pytest.log
tail -12 gpurun_out/ab_pytest.log
echo "== in-tree"; python bench.py --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>gpurun_out/ab_bench_tree.err | tee gpurun_out/ab_bench_tree.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
for lib in "$@"; do
  echo "== $lib"
  if [[ "$lib" == *prof* ]]; then
    GEOLDM_B200_LIB=$PWD/$lib python scripts/tc16_stats.py 2>&1 | tee gpurun_out/ab_stats_$(basename $lib .so).log
  else
    GEOLDM_B200_LIB=$PWD/$lib python bench.py --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>gpurun_out/ab_bench_$(basename $lib .so).err | tee gpurun_out/ab_bench_$(basename $lib .so).json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
  fi
done

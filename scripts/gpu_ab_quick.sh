#!/bin/bash
# A/B without the test suite: bench step time for the in-tree library and each variant, stats for *prof* variants.
mkdir -p gpurun_out
run() { GEOLDM_B200_LIB=$1 python bench.py --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>gpurun_out/abq.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value'],1), 'mol/s', round(d['ms_per_step'],4), 'ms/step  gcl', round(d['roofline']['kernel_ms'],4), d['clocks']['sm_mhz'])"; }
run $PWD/geoldm_b200/csrc/libgeoldm_b200.so in-tree
for lib in "$@"; do
  if [[ "$lib" == *prof* ]]; then GEOLDM_B200_LIB=$PWD/$lib python scripts/tc16_stats.py 2>&1 | tee gpurun_out/ab_stats_$(basename $lib .so).log
  else run $PWD/$lib $lib; fi
done
run $PWD/geoldm_b200/csrc/libgeoldm_b200.so in-tree-again

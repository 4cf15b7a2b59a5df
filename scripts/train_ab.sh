#!/bin/bash
# A/B of library variants (GEOLDM_B200_LIB) on the captured config-5 training step: scripts/train_ab.sh [variant.so ...]
mkdir -p gpurun_out
run() { GEOLDM_B200_LIB=$1 python scripts/train_graph_profile.py 2>/dev/null | grep -E "gpu_busy|gemm_tn|linear_small|edge_tail_bwd|edge_act_bwd" | tr '\n' ' ' | sed "s/^/$2: /"; echo; }
run $PWD/geoldm_b200/csrc/libgeoldm_b200.so in-tree
for lib in "$@"; do run $PWD/$lib $(basename $lib .so); done
run $PWD/geoldm_b200/csrc/libgeoldm_b200.so in-tree-again

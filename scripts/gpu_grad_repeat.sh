#!/bin/bash
# 12 repetitions of the config-5 gradient golden: run-to-run spread of the worst tensor (fp32 atomics in the backward kernels)
for r in 1 2 3 4 5 6 7 8 9 10 11 12; do timeout 200 python -m pytest tests/test_train_step.py -q -x -m gpu -s -k "qm9cond" 2>&1 | grep -o "worst grad vs fp64 reference [0-9.e-]* ([a-z_.0-9]*), single-element tensors [0-9.e-]*\|passed\|failed" | tr '\n' ' '; echo; done

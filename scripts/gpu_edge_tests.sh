#!/bin/bash
# GPU parity subset that exercises the edge kernels (quick check after a kernel change)
python -m pytest tests/test_gpu_parity.py -q -x -k "edge_kernels or qm9_forward or flag_variants or geom_forward or degenerate or full_size_properties" 2>&1 | tail -3

GEOLDM_B200_LIB=$PWD/ab/lib_sps2.so python -m pytest tests/test_gpu_parity.py -q -x -k "edge_kernels or qm9_forward or flag_variants or geom_forward or degenerate" 2>&1 | tail -3

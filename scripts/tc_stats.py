"""GEOLDM_TC_DEBUG=32 python scripts/tc_stats.py : where the MMA-issuing thread waits (edge GCL/EQUIV + dense launches)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from geoldm_b200 import _lib
from geoldm_b200.models import get_latent_diffusion
from geoldm_b200.packing import pack_molecules
L = _lib.lib()
dev = torch.device("cuda:0")
margs = bench.qm9_args("3xtf32")
info = {"atom_decoder": list("HCNOF"), "n_nodes": {5: 1}, "max_n_nodes": 29}
torch.manual_seed(0)
model, _, _ = get_latent_diffusion(margs, dev, info, None)
bench.tame_(model, 256)
nodes = bench.workload_nodes(1250)
batch = pack_molecules(nodes, dev)
dyn = model.dynamics
w, _keep = dyn.egnn.packed()
ccfg = dyn.egnn.c_config()
cb = batch.c_batch(128)
H = 256
pq = torch.randn(batch.n_node, 2 * H, device=dev)
xx = torch.randn(batch.n_node, 3, device=dev)
out = (C.c_ulonglong * 16)()
for name, fn, em, o in (("GCL", L.geoldm_edge_gcl, w.block[0].gcl[0].edge, torch.zeros(batch.n_node, H, device=dev)),
                        ("EQUIV", L.geoldm_edge_equiv, w.block[0].equiv, torch.zeros(batch.n_node, 3, device=dev))):
    for _ in range(3):
        _lib.check(fn(C.byref(ccfg), C.byref(em), C.byref(cb), _lib.ptr(pq), _lib.ptr(xx), _lib.ptr(xx), _lib.ptr(o), None), name)
    L.geoldm_tc_read_stats(out)
    n = 5
    for _ in range(n):
        _lib.check(fn(C.byref(ccfg), C.byref(em), C.byref(cb), _lib.ptr(pq), _lib.ptr(xx), _lib.ptr(xx), _lib.ptr(o), None), name)
    L.geoldm_tc_read_stats(out)
    tot, acc, a, wv, launches, tiles = [out[i] for i in range(6)]
    print(f"{name}: per tile-pair cycles: total {tot/tiles:.0f} | wait acc_empty {acc/tiles:.0f} | wait a_full {a/tiles:.0f} | "
          f"wait w_full {wv/tiles:.0f} | issue+other {(tot-acc-a-wv)/tiles:.0f}   ({launches} launches, {tiles/launches:.0f} tile-pairs per CTA)")
    print(f"   producer thread per tile: loads+wait a_empty {out[6]/tiles:.0f} | compute+STS {out[7]/tiles:.0f} | fence+arrive {out[8]/tiles:.0f} | tile metadata {out[9]/tiles:.0f}")

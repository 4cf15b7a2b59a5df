"""GEOLDM_TC_PROFILE=1 python -m geoldm_b200.build --force; python scripts/tc16_stats.py
Cycle breakdown of the fp16-split edge kernels (MMA thread, one producer thread, one epilogue thread of CTA 0)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from geoldm_b200 import _lib
from geoldm_b200.models import get_latent_diffusion
from geoldm_b200.packing import pack_molecules
L = _lib.lib()
dev = torch.device("cuda:0")
margs = bench.qm9_args("3xf16")
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
import os as _os
out = (C.c_ulonglong * 16)()
# same inputs as geoldm_egnn_forward feeds the kernels: per-edge squared distances and normalised differences
r_e = torch.empty(batch.n_edge, device=dev)
u_e = torch.empty(batch.n_edge, 4, device=dev)
_lib.check(L.geoldm_edge_dist(C.byref(cb), _lib.ptr(xx), _lib.ptr(r_e), _lib.ptr(u_e), 1.0, None), "edge_dist")
o_gcl, o_eq = torch.zeros(batch.n_node, H, device=dev), torch.zeros(batch.n_node, 3, device=dev)
gcl_e, eq_e = w.block[0].gcl[0].edge, w.block[0].equiv
runs = (("GCL", lambda: L.geoldm_edge_gcl_pre(C.byref(ccfg), C.byref(gcl_e), C.byref(cb), _lib.ptr(pq), 2 * H, _lib.ptr(r_e),
                                               _lib.ptr(r_e), _lib.ptr(o_gcl), None)),
        ("EQUIV", lambda: L.geoldm_edge_equiv_pre(C.byref(ccfg), C.byref(eq_e), C.byref(cb), _lib.ptr(pq), 2 * H, _lib.ptr(r_e),
                                                   _lib.ptr(r_e), _lib.ptr(u_e), _lib.ptr(o_eq), None)))
for name, fn in runs:
    for _ in range(3):
        _lib.check(fn(), name)
    L.geoldm_tc16_read_stats(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(fn(), name)
    e1.record()
    L.geoldm_tc16_read_stats(out)
    print(f"{name}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch (events, 5 back-to-back launches)")
    v = [float(out[i]) for i in range(16)]
    t = v[5]
    print(f"{name}: per tile-pair cycles: MMA thread total {v[0]/t:.0f} | wait acc_empty {v[1]/t:.0f} | wait a_full {v[2]/t:.0f} | "
          f"wait w {v[3]/t:.0f} | issue+other {(v[0]-v[1]-v[2]-v[3])/t:.0f}   ({v[4]:.0f} launches, {t/v[4]:.0f} tile-pairs per CTA)")
    print(f"   producer thread: wait a_empty(+loads) {v[6]/t:.0f} | compute+store {v[7]/t:.0f} | fence+arrive {v[8]/t:.0f} | metadata {v[9]/t:.0f}")
    print(f"   CTA 0 entry->exit {v[15]/v[4]:.0f} cycles per launch = {v[15]/v[4]/1.965e3:.1f} us")
    print(f"   epilogue thread: wait acc_full {v[10]/t:.0f} | pass 1 {v[11]/t:.0f} | dot exchange+gate {v[12]/t:.0f} | pass 2 {v[13]/t:.0f} | metadata {v[14]/t:.0f}")

# ---- node-level (dense) launches of one block ---------------------------------------------------------------------
N = batch.n_node
hbuf = torch.randn(N, H, device=dev)
agg = torch.randn(N, H, device=dev)
t1 = torch.empty(N, H, device=dev)
h2 = torch.empty(N, H, device=dev)
pq4 = torch.empty(N, 4 * H, device=dev)
g0 = w.block[0].gcl[0]
cases = (("node1 K=512 N=256", lambda: L.geoldm_linear_tc(H, 16, _lib.ptr(hbuf), H, _lib.ptr(agg), H, 1.0, g0.tc_pack_node1, 1, g0.node_b1, None, 1, _lib.ptr(t1), N, None)),
         ("node2 K=256 N=256", lambda: L.geoldm_linear_tc(H, 16, _lib.ptr(t1), H, None, 0, 1.0, g0.tc_pack_node2, 1, g0.node_b2, _lib.ptr(hbuf), 2, _lib.ptr(h2), N, None)),
         ("pq4   K=256 N=1024", lambda: L.geoldm_linear_tc(H, 16, _lib.ptr(hbuf), H, None, 0, 1.0, w.block[0].tc_pack_pq4, 4, w.block[0].pq4_b, None, 0, _lib.ptr(pq4), N, None)))
for name, fn in cases:
    for _ in range(3):
        _lib.check(fn(), name)
    L.geoldm_tc16_read_stats(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(fn(), name)
    e1.record()
    L.geoldm_tc16_read_stats(out)
    v = [float(out[i]) for i in range(16)]
    t = v[5]
    print(f"{name}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us/launch; CTA 0: {t/v[4]:.0f} tile-pairs, MMA thread total/launch {v[0]/v[4]:.0f} cycles; per tile-pair: "
          f"wait acc_empty {v[1]/t:.0f} | wait a_full {v[2]/t:.0f} | wait w {v[3]/t:.0f} | issue {(v[0]-v[1]-v[2]-v[3])/t:.0f}")
    print(f"   CTA 0 entry->exit {v[15]/v[4]:.0f} cycles per launch = {v[15]/v[4]/1.965e3:.1f} us")
    print(f"   producer: wait a_empty(+loads) {v[6]/t:.0f} | compute+store {v[7]/t:.0f} | fence+arrive {v[8]/t:.0f} | metadata {v[9]/t:.0f};"
          f"  epilogue: wait acc_full {v[10]/t:.0f} | metadata {v[14]/t:.0f}")

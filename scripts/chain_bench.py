"""Fused node-chain kernel vs the three dense launches it replaces (us per launch, N = nodes of the 1250-molecule workload)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geoldm_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
H = 256
N = int(os.environ.get("ROWS", 22576))
g = torch.Generator().manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).to(dev)
h, agg = mk(N, H), mk(N, H)
w1, b1, w2, b2 = mk(H, 2 * H) / 22, mk(H), mk(H, H) / 16, mk(H)
def pack(w, n_out, k):
    p = torch.empty(L.geoldm_tc_pack16_bytes(H, n_out, k), dtype=torch.uint8, device=dev)
    _lib.check(L.geoldm_tc_pack16(H, _lib.ptr(w), n_out, k, _lib.ptr(p), None), "pack"); return p
p1, p2 = pack(w1, H, 2 * H), pack(w2, H, H)
t1, h2 = torch.empty(N, H, device=dev), torch.empty(N, H, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for nb in (1, 2, 4):
    w3, b3 = mk(nb * H, H) / 16, mk(nb * H)
    p3 = pack(w3, nb * H, H)
    pq = torch.empty(N, nb * H, device=dev)
    chain = lambda: _lib.check(L.geoldm_node_chain(H, _lib.ptr(h), _lib.ptr(agg), 100.0, _lib.ptr(p1), _lib.ptr(b1), _lib.ptr(p2), _lib.ptr(b2),
                                                   _lib.ptr(p3), _lib.ptr(b3), nb, _lib.ptr(h2), _lib.ptr(pq), None, N, None), "chain")
    def three():
        _lib.check(L.geoldm_linear_tc(H, 16, _lib.ptr(h), H, _lib.ptr(agg), H, 100.0, _lib.ptr(p1), 1, _lib.ptr(b1), None, 1, _lib.ptr(t1), N, None), "n1")
        _lib.check(L.geoldm_linear_tc(H, 16, _lib.ptr(t1), H, None, 0, 1.0, _lib.ptr(p2), 1, _lib.ptr(b2), _lib.ptr(h), 2, _lib.ptr(h2), N, None), "n2")
        _lib.check(L.geoldm_linear_tc(H, 16, _lib.ptr(h2), H, None, 0, 1.0, _lib.ptr(p3), nb, _lib.ptr(b3), None, 0, _lib.ptr(pq), N, None), "pq")
    print(f"rows {N} n_pb {nb}: fused chain {timeit(chain):.1f} us, three dense launches {timeit(three):.1f} us")
    if os.environ.get("CHAIN_STATS"):
        out = (C.c_ulonglong * 16)()
        L.geoldm_tc16_read_stats(out)
        for _ in range(5): chain()
        L.geoldm_tc16_read_stats(out)
        v = [float(out[i]) for i in range(16)]; t = v[5]
        print(f"   MMA thread per tile-pair (CTA 0, {t / v[4]:.0f} tile-pairs/launch): total {v[0]/t:.0f} | phase1 {v[8]/t:.0f} (wait a_full {v[2]/t:.0f}) | "
              f"phase2 {v[9]/t:.0f} (wait image {v[6]/t:.0f}) | phase3 {v[10]/t:.0f} (wait image {v[7]/t:.0f}) | wait acc_empty {v[1]/t:.0f} | wait w {v[3]/t:.0f}")

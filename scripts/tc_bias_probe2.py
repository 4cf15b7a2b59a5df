"""Bin the residual signed relative error of the 3xTF32 GEMM by the output's mantissa / magnitude."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geoldm_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
H, terms = 256, 3
g = torch.Generator().manual_seed(3)
n_src, n_rows = 8192, 128 * 64
a = torch.nn.functional.silu(torch.randn(n_src, 2 * H, generator=g) * 2)
w = (torch.rand(H, H, generator=g) * 2 - 1) / np.sqrt(H)
src = torch.arange(n_rows, dtype=torch.int32)
tile_row = torch.arange(0, n_rows + 1, 128, dtype=torch.int32)
ad, wd, sd_, td = a.to(dev), w.to(dev), src.to(dev), tile_row.to(dev)
pack = torch.empty(L.geoldm_tc_pack_bytes(H, H, H), dtype=torch.uint8, device=dev)
_lib.check(L.geoldm_tc_pack(H, _lib.ptr(wd), H, H, _lib.ptr(pack), None), "pack")
out = torch.zeros(n_rows, H, device=dev)
_lib.check(L.geoldm_tc_selftest(H, terms, _lib.ptr(ad), _lib.ptr(sd_), _lib.ptr(td), n_rows // 128, n_rows,
                                _lib.ptr(pack), _lib.ptr(out), None), "selftest")
torch.cuda.synchronize()
ref = (a[:n_rows, :H].double() @ w.double().T).numpy().ravel()
o = out.cpu().double().numpy().ravel() / (1.0 + 1.6e-8 * 96)      # undo the built-in compensation
err = (o - ref) * np.sign(ref)                                    # signed toward/away from zero, absolute
mant, expo = np.frexp(np.abs(ref))                                # mant in [0.5, 1)
ulp = np.ldexp(1.0, expo - 24)                                    # fp32 ulp of |ref|
e_ulp = err / ulp
print("overall: mean err in ulp", e_ulp.mean(), "std", e_ulp.std(), " (n_adds=96)")
for lo in np.arange(0.5, 1.0, 0.0625):
    m = (mant >= lo) & (mant < lo + 0.0625) & (np.abs(ref) > 0.05)
    print(f"mant [{lo:.4f},{lo+0.0625:.4f}): n={m.sum():7d} mean {e_ulp[m].mean():+.3f} ulp  std {e_ulp[m].std():.3f} ulp | rel mean {(err[m]/np.abs(ref[m])).mean():+.3e}")
for lo, hi in ((0, 0.01), (0.01, 0.05), (0.05, 0.2), (0.2, 0.5), (0.5, 1), (1, 10)):
    m = (np.abs(ref) >= lo) & (np.abs(ref) < hi)
    if m.sum():
        print(f"|ref| in [{lo},{hi}): n={m.sum():7d} mean {e_ulp[m].mean():+.3f} ulp std {e_ulp[m].std():.3f} | abs mean {err[m].mean():+.3e} std {err[m].std():.3e}")

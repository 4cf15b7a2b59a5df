"""Measure the signed relative error of the tcgen05 split GEMMs (selftest kernel; terms 3 = 3xTF32, 1 = TF32,
16 = 3xF16) vs fp64: mean (residual bias after the round-toward-zero compensation) and spread."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geoldm_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
for H in (64, 128, 192, 256):
    for terms in (16, 3, 1):
        for dist in ("gauss", "silu"):
            g = torch.Generator().manual_seed(H)
            n_src, n_rows = 2048, 128 * 16
            a = torch.randn(n_src, 2 * H, generator=g)
            if dist == "silu":
                a = torch.nn.functional.silu(a * 2)
            w = (torch.rand(H, H, generator=g) * 2 - 1) / np.sqrt(H)
            src = torch.arange(n_rows, dtype=torch.int32)
            tile_row = torch.arange(0, n_rows + 1, 128, dtype=torch.int32)
            ad, wd, sd_, td = a.to(dev), w.to(dev), src.to(dev), tile_row.to(dev)
            if terms == 16:
                pack = torch.empty(L.geoldm_tc_pack16_bytes(H, H, H), dtype=torch.uint8, device=dev)
                _lib.check(L.geoldm_tc_pack16(H, _lib.ptr(wd), H, H, _lib.ptr(pack), None), "pack16")
            else:
                pack = torch.empty(L.geoldm_tc_pack_bytes(H, H, H), dtype=torch.uint8, device=dev)
                _lib.check(L.geoldm_tc_pack(H, _lib.ptr(wd), H, H, _lib.ptr(pack), None), "pack")
            out = torch.zeros(n_rows, H, device=dev)
            _lib.check(L.geoldm_tc_selftest(H, terms, _lib.ptr(ad), _lib.ptr(sd_), _lib.ptr(td), n_rows // 128, n_rows,
                                            _lib.ptr(pack), _lib.ptr(out), None), "selftest")
            torch.cuda.synchronize()
            ref = a[:n_rows, :H].double() @ w.double().T
            o = out.cpu().double()
            big = ref.abs() > 0.25 * ref.abs().max()
            rel = ((o - ref) / ref.abs())[big] * torch.sign(ref[big])     # negative = shrunk toward zero
            ref32 = (a[:n_rows, :H] @ w.T).double()
            rel32 = ((ref32 - ref) / ref.abs())[big]
            n_adds = (H // 16) * 3 if terms == 16 else (H // 8) * terms
            print(f"H={H} terms={terms} {dist}: adds={n_adds} signed rel err mean {rel.mean():+.3e} std {rel.std():.3e} "
                  f"| per add {rel.mean()/n_adds:+.3e} | max/max {float((o-ref).abs().max()/ref.abs().max()):.2e} "
                  f"| torch fp32 cpu: mean {rel32.mean():+.2e} std {rel32.std():.2e}")

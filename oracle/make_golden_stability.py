"""tests/golden/stability.npz: outputs of the UNMODIFIED reference's check_stability / analyze_stability_for_molecules
(qm9/analyze.py) on synthetic molecules whose neighbour distances straddle the bond thresholds.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_stability.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import make_golden as G            # noqa: E402


def synth_molecule(rng, n, n_types, type_p):
    """Random tree: every new atom sits 0.85-1.75 A from a random earlier atom (so that many pairs fall next to a
    single/double/triple threshold), plus a little jitter."""
    pos = np.zeros((n, 3), dtype=np.float64)
    for k in range(1, n):
        parent = rng.integers(0, k)
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        pos[k] = pos[parent] + v * rng.uniform(0.85, 1.75)
    types = rng.choice(n_types, size=n, p=type_p)
    return pos.astype(np.float32), types.astype(np.int64)


def known_molecules():
    """Small real molecules (idealised geometries, Angstrom): all stable under the reference's rules."""
    t = 1.09 / np.sqrt(3.0)
    return [
        (np.array([[0, 0, 0], [t, t, t], [t, -t, -t], [-t, t, -t], [-t, -t, t]]), [1, 0, 0, 0, 0]),          # CH4
        (np.array([[0, 0, 0], [0.96, 0, 0], [-0.24, 0.93, 0]]), [3, 0, 0]),                                       # H2O
        (np.array([[0, 0, 0], [0.74, 0, 0]]), [0, 0]),                                                            # H2
        (np.array([[0, 0, 0], [1.16, 0, 0], [-1.16, 0, 0]]), [1, 3, 3]),                                          # CO2
        (np.array([[0, 0, 0], [1.156, 0, 0], [-1.06, 0, 0]]), [1, 2, 0]),                                         # HCN
        (np.array([[0, 0, 0], [0.94, 0.38, 0], [-0.47, 0.38, 0.81], [-0.47, 0.38, -0.81]]), [2, 0, 0, 0]),        # NH3
        (np.array([[0, 0, 0], [0.92, 0, 0]]), [4, 0]),                                                            # HF
    ]


def main():
    dc, qm, qs = G.import_reference()
    import qm9.analyze as ref_an
    rng = np.random.default_rng(2024)
    arrays = {}
    for ds, n_lo, n_hi, count in (("qm9", 3, 29, 40), ("geom", 8, 70, 12)):
        info = dc.get_dataset_info(ds, False)
        T = len(info["atom_decoder"])
        p = np.ones(T)
        if ds == "qm9":
            p = np.array([0.5, 0.3, 0.08, 0.1, 0.02])
        else:
            p[[0, 2, 3, 4]] = [12, 8, 3, 3]
        p = p / p.sum()
        n_max = n_hi
        X = np.zeros((count, n_max, 3), dtype=np.float32)
        A = np.zeros((count, n_max), dtype=np.int64)
        NN = np.zeros(count, dtype=np.int64)
        res = np.zeros((count, 3), dtype=np.int64)
        for m in range(count):
            n = int(rng.integers(n_lo, n_hi + 1))
            pos, types = synth_molecule(rng, n, T, p)
            if ds == "qm9" and m < len(known_molecules()):
                kp, kt = known_molecules()[m]
                Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
                pos, types, n = (kp @ Q.T + rng.normal(size=3)).astype(np.float32), np.array(kt, dtype=np.int64), len(kt)
            if m % 5 == 0 and n >= 4 and m >= 7:           # exact threshold hits: put atom 1 at (threshold/100) from atom 0
                pos[1] = pos[0] + np.array([rng.choice([1.19, 1.23, 1.39, 1.64, 1.53]), 0, 0], dtype=np.float32)
            X[m, :n], A[m, :n], NN[m] = pos, types, n
            # alternate the two input kinds the reference uses: torch tensors (analyze) and numpy arrays (sample_chain)
            if m % 2 == 0:
                out = ref_an.check_stability(torch.from_numpy(pos), torch.from_numpy(types), info)
            else:
                out = ref_an.check_stability(pos, types, info)
            res[m] = [int(out[0]), int(out[1]), int(out[2])]
        one_hot = torch.nn.functional.one_hot(torch.from_numpy(A), T) * (torch.arange(n_max)[None, :, None] < torch.from_numpy(NN)[:, None, None])
        node_mask = (torch.arange(n_max)[None, :] < torch.from_numpy(NN)[:, None]).float()
        ref_an.use_rdkit = False
        validity, _ = ref_an.analyze_stability_for_molecules(
            {"one_hot": one_hot, "x": torch.from_numpy(X), "node_mask": node_mask}, info)
        arrays.update({f"{ds}_x": X, f"{ds}_types": A, f"{ds}_n": NN, f"{ds}_res": res,
                       f"{ds}_validity": np.array([validity["mol_stable"], validity["atm_stable"]])})
        print(ds, "stable molecules", res[:, 0].sum(), "of", count, "stable atoms", res[:, 1].sum(), "of", res[:, 2].sum())
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "stability.npz")
    np.savez_compressed(out, **arrays)
    print("wrote", out)


if __name__ == "__main__":
    main()

"""tests/golden/chain_*.npz: EnLatentDiffusion.sample_chain and qm9.sampling.sample_chain / sample_sweep_conditional of the
UNMODIFIED reference on a small tamed model.  Noise comes from torch's global CPU generator after
torch.manual_seed(seed); the tests rebuild the same draws in call order (x block, then h block, T+2 times).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_chain.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import geoldm_oracle as O          # noqa: E402
from oracle import make_golden as G            # noqa: E402


def main():
    refmods = G.import_reference()
    dc, qm, qs = refmods
    torch.set_num_threads(os.cpu_count() or 1)
    c = O.OracleConfig(nf=64, n_layers=2)
    m, a, info, _, _ = G.build_reference(c, "qm9", 9, True, refmods)
    # ---- model-level sample_chain, two molecules of different size, 20 frames ------------------------
    nodes = [7, 12]
    nm, em = O.build_masks(nodes, 12)
    torch.manual_seed(78)
    with torch.no_grad():
        chain = m.sample_chain(2, 12, nm, em, None, keep_frames=20)
    G.save("chain_model", c, "qm9", 9, True, nodes=np.array(nodes), torch_seed=np.array([78]), keep_frames=np.array([20]),
           chain=chain)
    # ---- qm9/sampling.py:sample_chain (19 atoms, 100 frames + 10 repeats, stability of the last frame) ------------
    torch.manual_seed(79)
    with torch.no_grad():
        one_hot, charges, x = qs.sample_chain(a, "cpu", m, 1, info)
    G.save("chain_sampling", c, "qm9", 9, True, torch_seed=np.array([79]), one_hot=one_hot, charges=charges, x=x)
    # ---- sample_sweep_conditional on a conditional small model ------------------------------------------------
    cc = O.OracleConfig(nf=64, n_layers=2, context_node_nf=1, include_charges=False, normalize_factors=(1.0, 8.0, 1.0))
    mc, ac, infoc, _, _ = G.build_reference(cc, "qm9_second_half", 9, True, refmods)
    prop = types.SimpleNamespace(distributions={"alpha": {9: {"params": (40.0, 90.0)}}},
                                 normalizer={"alpha": {"mean": torch.tensor(75.0), "mad": torch.tensor(6.0)}})
    torch.manual_seed(80)
    with torch.no_grad():
        one_hot, charges, x, node_mask = qs.sample_sweep_conditional(ac, "cpu", mc, infoc, prop, n_nodes=9, n_frames=6)
    G.save("chain_sweep", cc, "qm9_second_half", 9, True, torch_seed=np.array([80]), one_hot=one_hot, x=x,
           node_mask=node_mask)


if __name__ == "__main__":
    main()

"""CPU-only tests: C-ABI exports, ragged packing, sharding logic, Philox known answers."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from geoldm_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "geoldm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(geoldm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    handle = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/geoldm_b200.h but not exported"
    from geoldm_b200 import _lib
    assert set(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ declared
    assert _lib.lib().geoldm_abi_version() == 4


def test_product_has_no_cpu_fallback():
    from geoldm_b200 import _lib
    from geoldm_b200.dynamics import EGNN_dynamics_QM9
    dyn = EGNN_dynamics_QM9(in_node_nf=2, context_node_nf=0, n_dims=3, hidden_nf=32, n_layers=1, attention=True,
                            tanh=True, norm_constant=1, inv_sublayers=1, normalization_factor=1)
    nm = torch.ones(1, 4, 1)
    em = (torch.ones(4, 4) - torch.eye(4)).reshape(-1, 1)
    with pytest.raises(_lib.GeoldmError):
        dyn._forward(torch.tensor([[0.5]]), torch.zeros(1, 4, 4), nm, em, None)


def test_product_never_imports_oracle_or_reference():
    bad = re.compile(r"^\s*(from|import)\s+(oracle|egnn\b|equivariant_diffusion|qm9\b)", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "geoldm_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f
                assert "/root/reference" not in src, f


def test_state_dict_layout_matches_reference_keys():
    """139 dynamics tensors for L=9, S=1, plus decoder/encoder/gamma/buffers (304 entries)."""
    from oracle import geoldm_oracle as O
    from tests.helpers import make_args
    from geoldm_b200.models import get_latent_diffusion
    cfg = O.QM9_CFG
    info = {"atom_decoder": list(range(5)), "n_nodes": {5: 1}, "max_n_nodes": 29}
    model, nodes_dist, _ = get_latent_diffusion(make_args(cfg), "cpu", info, None)
    sd = model.state_dict()
    assert len(sd) == 304
    dyn = [k for k in sd if k.startswith("dynamics.egnn.")]
    assert len(dyn) == 139
    ref = O.make_state_dict(cfg, 0)
    for k, v in ref.items():
        assert k in sd and sd[k].shape == v.shape, k
    assert sum(p.numel() for p in model.dynamics.parameters()) == 5337355
    assert torch.equal(sd["gamma.gamma"], ref["gamma.gamma"])


def test_packing_tables():
    from geoldm_b200.packing import pack_from_masks, pack_molecules
    b = pack_molecules([3, 1, 4], "cpu", n_max=5)
    assert b.n_node == 8 and b.n_edge == 3 * 2 + 0 + 4 * 3
    assert b.mol_off.tolist() == [0, 3, 4, 8]
    assert b.node_mol.tolist() == [0, 0, 0, 1, 2, 2, 2, 2]
    assert b.node_src.tolist() == [0, 1, 2, 5, 10, 11, 12, 13]
    ei, ej = b.edge_i.tolist(), b.edge_j.tolist()
    assert ei[:6] == [0, 0, 1, 1, 2, 2] and ej[:6] == [1, 2, 0, 2, 0, 1]
    assert all(i != j for i, j in zip(ei, ej)) and ei == sorted(ei)
    cb = b.c_batch(4)
    assert cb.n_tile == 5 and cb.tile_m == 4
    # masks incl. a non-prefix mask
    nm = torch.tensor([[1, 1, 0, 1], [0, 1, 1, 0]], dtype=torch.float32).unsqueeze(2)
    em = nm.unsqueeze(1).squeeze(3) * nm.squeeze(2).unsqueeze(2) * (1 - torch.eye(4)).unsqueeze(0)
    b2 = pack_from_masks(nm, em.reshape(-1, 1))
    assert b2.n_nodes.tolist() == [3, 2] and b2.node_src.tolist() == [0, 1, 3, 5, 6]
    with pytest.raises(ValueError):
        pack_from_masks(nm, torch.ones(2 * 16, 1))
    with pytest.raises(ValueError):
        pack_from_masks(torch.zeros(1, 4, 1))


def test_balance_shards():
    from geoldm_b200.packing import balance_shards
    rng = np.random.default_rng(0)
    n = rng.integers(3, 30, size=1000)
    shards = balance_shards(n, 8)
    assert sorted(np.concatenate(shards).tolist()) == list(range(1000))
    loads = [int((n[s] * (n[s] - 1)).sum()) for s in shards]
    assert max(loads) / min(loads) < 1.01
    n = rng.integers(3, 30, size=5000)                       # large jobs: the vectorised deal
    shards = balance_shards(n, 8)
    assert sorted(np.concatenate(shards).tolist()) == list(range(5000))
    assert all(np.all(np.diff(s) > 0) for s in shards)
    loads = [int((n[s] * (n[s] - 1)).sum()) for s in shards]
    assert max(loads) / min(loads) < 1.005


def test_balance_shards_equal_counts():
    """Training batches: equal molecule counts per rank (per-rank mean losses are averaged), edge counts within 2 %."""
    from geoldm_b200.packing import balance_shards_equal
    rng = np.random.default_rng(1)
    for world in (2, 4, 8):
        n = rng.integers(3, 30, size=64 * world)
        shards = balance_shards_equal(n, world)
        assert all(len(s) == 64 for s in shards) and all(np.all(np.diff(s) > 0) for s in shards)
        assert sorted(np.concatenate(shards).tolist()) == list(range(64 * world))
        loads = [int((n[s] * (n[s] - 1)).sum()) for s in shards]
        assert max(loads) / min(loads) < 1.02, (world, loads)
    with pytest.raises(ValueError):
        balance_shards_equal(np.arange(3, 13), 4)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    from tests.philox_ref import philox4x32_10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_step_table_matches_oracle_coefficients():
    from oracle import geoldm_oracle as O
    from tests.helpers import make_args
    from geoldm_b200.models import get_latent_diffusion
    cfg = O.OracleConfig(nf=32, n_layers=1)
    info = {"atom_decoder": list(range(5)), "n_nodes": {5: 1}, "max_n_nodes": 29}
    model, _, _ = get_latent_diffusion(make_args(cfg), "cpu", info, None)
    table = model.step_table("cpu")
    gamma = torch.from_numpy(O.noise_schedule_gamma(cfg))
    for s in (0, 1, 499, 998, 999):
        a, c, n = O.step_coefficients(gamma, 1000, s)
        assert table[s, 0] == a and table[s, 1] == c and table[s, 2] == n
        assert table[s, 3] == torch.tensor(float(s + 1)) / 1000
    assert table.shape == (1001, 4) and table[1000, 3] == 0


def test_model_deepcopy_for_ema():
    """main_qm9.py:227-231 clones the model for the EMA weights; cached device images must not block that."""
    import copy
    from tests.helpers import build_cuda_model
    from oracle import geoldm_oracle as O
    cfg = O.OracleConfig(nf=32, n_layers=1)
    model = build_cuda_model(cfg, O.make_state_dict(cfg, 0), device="cpu")
    clone = copy.deepcopy(model)
    a, b = model.state_dict(), clone.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert all(p.data_ptr() != q.data_ptr() for p, q in zip(model.parameters(), clone.parameters()) if p.numel())


def test_product_distribution_nodes_matches_reference_draws():
    """a17: the PRODUCT class geoldm_b200.models.DistributionNodes (not the oracle's function) reproduces the reference's
    draws for torch.manual_seed(0) (fixture written by oracle/make_golden.py from qm9/models.py:178-215) and its
    log-probabilities."""
    import numpy as np
    from geoldm_b200.histograms import QM9_WITH_H_N_NODES
    from geoldm_b200.models import DistributionNodes
    from tests.helpers import GOLDEN
    want = np.load(os.path.join(GOLDEN, "nodes_dist_qm9_seed0.npz"))["draws"]
    torch.manual_seed(0)
    nd = DistributionNodes(QM9_WITH_H_N_NODES)
    got = nd.sample(64)
    assert np.array_equal(got.numpy(), want)
    p = np.array(list(QM9_WITH_H_N_NODES.values()), dtype=np.float64)
    lp = nd.log_prob(torch.tensor([29, 9, 18]))
    idx = [list(QM9_WITH_H_N_NODES.keys()).index(k) for k in (29, 9, 18)]
    assert np.allclose(lp.numpy(), np.log(p[idx] / p.sum()), rtol=1e-6)


def test_gradient_buckets_cover_every_trainable_parameter():
    """training.gradient_buckets: latent model (decoder + denoiser; encoder too when the first stage is trainable) and a
    stand-alone first-stage EnHierarchicalVAE ('encoder.*' / 'decoder.*' names) - no trainable parameter is left out,
    so multi-rank training can never skip the all-reduce silently."""
    from oracle import geoldm_oracle as O
    from tests.helpers import make_args
    from geoldm_b200.models import get_autoencoder, get_latent_diffusion
    from geoldm_b200.training import gradient_buckets
    cfg = O.OracleConfig(nf=32, n_layers=1)
    info = {"atom_decoder": list(range(5)), "n_nodes": {5: 1}, "max_n_nodes": 29}
    for build, kw in ((get_latent_diffusion, dict(trainable_ae=True)), (get_latent_diffusion, dict(trainable_ae=False)),
                      (get_autoencoder, dict(trainable_ae=True))):
        model = build(make_args(cfg, "fp32", **kw), "cpu", info, None)[0]
        want = {id(p) for p in model.parameters() if p.requires_grad}
        got = [id(p) for b in gradient_buckets(model) for p in b]
        assert len(got) == len(set(got)) and set(got) == want and len(want) > 0, (build.__name__, kw)

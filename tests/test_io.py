"""the reference's on-disk formats.  The checkpoint directory and the xyz files under tests/golden/
were written by the unmodified reference's own code (oracle/make_golden_io.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import geoldm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INFO = {"name": "qm9", "atom_decoder": ['H', 'C', 'N', 'O', 'F'], "atom_encoder": {'H': 0, 'C': 1, 'N': 2, 'O': 3, 'F': 4},
        "n_nodes": {5: 3, 9: 4, 19: 2}, "max_n_nodes": 29}


def test_reference_checkpoint_loads_strictly():
    from geoldm_b200.io import load_checkpoint
    model, nodes_dist, prop_dist, args = load_checkpoint(os.path.join(GOLDEN, "ckpt_small"), "cpu", INFO, mma_mode="fp32")
    assert args.nf == 32 and args.n_layers == 1 and prop_dist is None and not model.training
    ref = torch.load(os.path.join(GOLDEN, "ckpt_small", "generative_model_ema.npy"), map_location="cpu")
    own = model.state_dict()
    assert own.keys() == ref.keys()
    assert all(torch.equal(own[k], ref[k]) for k in ref)


def test_xyz_writer_is_byte_identical(tmp_path):
    from geoldm_b200.io import load_molecule_xyz, save_xyz_file
    g = json.load(open(os.path.join(GOLDEN, "xyz_golden.json")))
    one_hot, x, nm = torch.tensor(g["one_hot"]), torch.tensor(g["x"]), torch.tensor(g["node_mask"])
    d = str(tmp_path) + "/"
    save_xyz_file(d, one_hot, None, x, INFO, id_from=7, name="molecule", node_mask=nm)
    save_xyz_file(d, one_hot[:1], None, x[:1], INFO, id_from=0, name="full")
    assert sorted(os.listdir(d)) == sorted(g["files"])
    for fn, text in g["files"].items():
        assert open(os.path.join(d, fn)).read() == text, fn
    pos, oh, ch = load_molecule_xyz(os.path.join(d, "molecule_008.txt"), INFO)
    n = int(nm[1].sum())
    assert pos.shape == (n, 3) and torch.allclose(pos, x[1, :n], atol=1e-8) and torch.equal(oh, one_hot[1, :n])


@pytest.mark.gpu
def test_checkpoint_forward_and_analyze_on_gpu(tmp_path):
    import argparse
    from geoldm_b200.io import analyze_and_save, load_checkpoint
    path = os.path.join(GOLDEN, "ckpt_small")
    model, nodes_dist, prop_dist, args = load_checkpoint(path, "cuda", INFO, mma_mode="fp32")
    f = np.load(os.path.join(path, "forward.npz"))
    nm, em = O.build_masks(f["nodes"].tolist(), f["z"].shape[1])
    with torch.no_grad():
        out = model.dynamics._forward(torch.from_numpy(f["t"]).cuda(), torch.from_numpy(f["z"]).cuda(), nm.cuda(),
                                      em.cuda(), None)
    err = O.err_metric(out.cpu(), torch.from_numpy(f["out"]))
    print(f"[io] forward of the loaded reference checkpoint: err {err:.2e}")
    assert err < 2e-5
    # eval_analyze.py flow: sample, write xyz, stability (T shortened: the step count is a model attribute)
    model.T = 20
    model.gamma = type(model.gamma)('polynomial_2', timesteps=20, precision=1e-5).cuda()
    eval_args = argparse.Namespace(model_path=str(tmp_path))
    validity, rdkit = analyze_and_save(args, eval_args, "cuda", model, nodes_dist, None, INFO, n_samples=6, batch_size=3,
                                       save_to_xyz=True)
    assert rdkit is None and 0.0 <= validity["atm_stable"] <= 1.0 and 0.0 <= validity["mol_stable"] <= 1.0
    files = sorted(os.listdir(os.path.join(str(tmp_path), "eval", "analyzed_molecules")))
    assert files == ["molecule_%03d.txt" % i for i in range(6)]

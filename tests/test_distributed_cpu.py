"""world_size-2 gloo test of the molecule-sharding / final-gather logic (host side of SURVEY §8e) on CPU.
The CUDA sampler is replaced by a deterministic CPU stand-in keyed by (global molecule id, seed)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def fake_sample(args, device, model, dataset_info, nodesxsample=None, context=None, fix_noise=False, seed=0,
                mol_ids=None):
    n_max = dataset_info["max_n_nodes"]
    bs = len(nodesxsample)
    node_mask = (torch.arange(n_max).unsqueeze(0) < nodesxsample.unsqueeze(1)).float().unsqueeze(2)
    x = torch.zeros(bs, n_max, 3)
    one_hot = torch.zeros(bs, n_max, 5, dtype=torch.int64)
    charges = torch.zeros(bs, n_max, 1, dtype=torch.int64)
    for k, gid in enumerate(mol_ids):
        g = torch.Generator().manual_seed(int(gid) * 1000 + seed)
        x[k] = torch.randn(n_max, 3, generator=g)
        one_hot[k, :, int(gid) % 5] = 1
        charges[k] = int(gid)
    return one_hot * node_mask.long(), charges * node_mask.long(), x * node_mask, node_mask


def _worker(rank, world, port, nodes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geoldm_b200.distributed import sample_sharded, shard_indices
    info = {"max_n_nodes": 29}
    res = sample_sharded(None, "cpu", None, info, torch.tensor(nodes), seed=3, sample_fn=fake_sample)
    mine = shard_indices(nodes, world, rank)
    torch.save({"res": res, "mine": mine}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_sampling_matches_single_rank(tmp_path):
    rng = np.random.default_rng(0)
    nodes = rng.integers(3, 30, size=37).tolist()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, nodes, str(tmp_path)), nprocs=2, join=True)
    info = {"max_n_nodes": 29}
    ref = fake_sample(None, "cpu", None, info, nodesxsample=torch.tensor(nodes), seed=3, mol_ids=np.arange(len(nodes)))
    shards = []
    for r in range(2):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False)
        shards.append(set(d["mine"].tolist()))
        for got, want in zip(d["res"], ref):
            assert torch.equal(got, want)
    assert shards[0] | shards[1] == set(range(len(nodes))) and not (shards[0] & shards[1])


def test_sharded_sampling_with_an_empty_rank(tmp_path):
    """More ranks than molecules: the rank with an empty shard skips sampling and still takes part in the gather."""
    nodes = [17]
    port = 29950 + os.getpid() % 40
    mp.spawn(_worker, args=(2, port, nodes, str(tmp_path)), nprocs=2, join=True)
    info = {"max_n_nodes": 29}
    ref = fake_sample(None, "cpu", None, info, nodesxsample=torch.tensor(nodes), seed=3, mol_ids=np.arange(1))
    for r in range(2):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False)
        for got, want in zip(d["res"], ref):
            assert torch.equal(got, want)


# ---- training: gradient all-reduce over two ranks == single-rank gradients of the whole batch ----------------
def _patch_cpu_autograd():
    """Route the wrappers through train.py's library-op graph so the host logic can run on CPU in this test."""
    from geoldm_b200 import dynamics
    from geoldm_b200 import train as _train
    _train.allow_cpu_graph_check(True)       # test seam: library GEMMs stand in for the CUDA kernels (no GPU here)
    dynamics._EgnnWrapper._check_inputs = lambda self, xh, nm: None
    dynamics._EgnnWrapper._wants_grad = lambda self, xh, ctx: True


def _train_setup():
    import argparse
    from oracle import geoldm_oracle as O
    from tests.helpers import build_cuda_model, load_golden
    from geoldm_b200.histograms import HISTOGRAMS
    from geoldm_b200.models import DistributionNodes
    cfg, sd, A, meta = load_golden("train_small", encoder=True)
    model = build_cuda_model(cfg, sd, device="cpu", trainable_ae=True)
    # equal shard sizes so that mean-of-shard-means == mean of the whole batch: use molecules 0..3 -> 2 + 2
    sel = [0, 1, 2, 3]
    nm, em = O.build_masks(A["nodes"][sel].tolist(), A["x"].shape[1])
    data = dict(x=A["x"][sel], one_hot=A["one_hot"][sel], charges=A["charges"][sel], nm=nm,
                em=em.view(len(sel), -1), draws={k: A["train_" + k][sel] for k in ("eps_enc", "t_int", "eps_t")})
    args = argparse.Namespace(probabilistic_model="diffusion", lr=1e-3, clip_grad=True, ema_decay=0.0,
                              ode_regularization=0.0)
    return model, DistributionNodes(HISTOGRAMS["qm9"]), data, args


def _step(model, nodes_dist, data, args, idx, flat=False):
    from geoldm_b200.training import FlatGradBuckets, Queue, get_optim, train_step
    q = Queue()
    q.add(3000.0)
    optim = get_optim(args, model)
    n = data["x"].shape[1]
    pick = lambda t: t[idx]
    h = {"categorical": pick(data["one_hot"]), "integer": pick(data["charges"])}
    em = pick(data["em"]).reshape(-1, 1)
    buckets = FlatGradBuckets(model) if flat else None      # flat: all-reduce launched from backward hooks (overlap)
    nll, gn = train_step(args, model, optim, nodes_dist, pick(data["x"]), h, pick(data["nm"]), em, None,
                         gradnorm_queue=q, draws={k: pick(v) for k, v in data["draws"].items()}, buckets=buckets)
    return nll, float(gn)


def _train_worker(rank, world, port, out_dir, flat):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_cpu_autograd()
    model, nodes_dist, data, args = _train_setup()
    nll, gn = _step(model, nodes_dist, data, args, [2 * rank, 2 * rank + 1], flat)
    torch.save({"sd": model.state_dict(), "gn": gn, "nll": nll}, os.path.join(out_dir, f"t{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("flat", [False, True])
def test_two_rank_training_step_matches_single_rank(tmp_path, monkeypatch, flat):
    """flat=False: one all-reduce per bucket after backward; flat=True: FlatGradBuckets (gradients are views of flat
    buffers, all-reduces launched from post-accumulate hooks while backward runs)."""
    port = 29950 + os.getpid() % 40 + (50 if flat else 0)
    mp.spawn(_train_worker, args=(2, port, str(tmp_path), flat), nprocs=2, join=True)
    from geoldm_b200 import dynamics
    from geoldm_b200 import train as _train
    monkeypatch.setattr(_train, "_CPU_GRAPH_CHECK", True)
    monkeypatch.setattr(dynamics._EgnnWrapper, "_check_inputs", lambda self, xh, nm: None)
    monkeypatch.setattr(dynamics._EgnnWrapper, "_wants_grad", lambda self, xh, ctx: True)
    model, nodes_dist, data, args = _train_setup()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    nll, gn = _step(model, nodes_dist, data, args, [0, 1, 2, 3])
    single = model.state_dict()
    r0 = torch.load(os.path.join(tmp_path, "t0.pt"), weights_only=False)
    r1 = torch.load(os.path.join(tmp_path, "t1.pt"), weights_only=False)
    assert abs(r0["gn"] - gn) / gn < 1e-5 and abs(r1["gn"] - gn) / gn < 1e-5       # global norm after all-reduce
    assert abs(float(0.5 * (r0["nll"] + r1["nll"]) - nll)) < 1e-5 * abs(float(nll))
    moved = 0
    for k in single:
        assert torch.equal(r0["sd"][k], r1["sd"][k]), k                                 # ranks stay in lock-step
        step = (single[k] - before[k]).abs().max()
        if step > 0:
            moved += 1
            assert (r0["sd"][k] - single[k]).abs().max() <= 2e-2 * step + 1e-9, k        # same AdamW update
    assert moved > 50

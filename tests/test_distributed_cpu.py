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

#!/usr/bin/env python
"""bench.py — GeoLDM QM9 1000-step sampling throughput on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], weak scaling): 1250 synthetic QM9 molecules per GPU (10 000 at 8 GPUs),
atom counts drawn from the QM9 histogram (seed 0), EGNN_dynamics nf=256, 9 layers, latent_nf=1, T=1000,
random-init weights (no checkpoints offline).  One bench "step" = one pass of the hot path over the batch =
one ancestral sampling step (denoiser forward + z_s update), replayed from the captured CUDA graph.  A sampled
molecule costs 1002 such passes (1000 steps + p(x|z0) + the same-sized decoder EGNN), so
    value [molecules/s] = molecules_on_all_ranks / (1002 * ms_per_step)          (max over ranks)
and `e2e` measures one COMPLETE sampling job (all 1002 passes, decode, host<->device copies) through the
public API qm9-style `sample(args, device, model, dataset_info, nodesxsample=...)` with host inputs/outputs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MOLS_PER_GPU = 1250
FORWARDS_PER_MOLECULE = 1002
CPU_SAMPLE_MOLS = 64
METRIC = "GeoLDM molecules/sec (1000-step sampling)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def workload_nodes(n_total, seed=0):
    """Atom counts of the synthetic molecules: categorical over the QM9 histogram (dict order), numpy PCG64."""
    from geoldm_b200.histograms import QM9_WITH_H_N_NODES as hist
    keys = np.array(list(hist.keys()))
    p = np.array(list(hist.values()), dtype=np.float64)
    rng = np.random.default_rng(seed)
    return keys[rng.choice(len(keys), size=n_total, p=p / p.sum())]


def oracle_module():
    """The CPU oracle — imported ONLY by the cpu_baseline / --impl reference legs (it is the thing timed there)."""
    from oracle import geoldm_oracle as O
    return O


def qm9_args(mma_mode):
    """The argparse fields of the reference CLI for the QM9 GeoLDM config (main_qm9.py:23-133 defaults + README)."""
    return argparse.Namespace(
        ae_path=None, cuda=True, include_charges=True, context_node_nf=0, conditioning=[], latent_nf=1, nf=256,
        n_layers=9, attention=True, tanh=True, model="egnn_dynamics", norm_constant=1, inv_sublayers=1,
        sin_embedding=False, normalization_factor=1, aggregation_method="sum", kl_weight=0.01,
        normalize_factors=[1, 4, 10], condition_time=True, probabilistic_model="diffusion", diffusion_steps=1000,
        diffusion_noise_schedule="polynomial_2", diffusion_noise_precision=1e-5, diffusion_loss_type="l2",
        trainable_ae=False, ema_decay=0.999, dataset="qm9", remove_h=False, mma_mode=mma_mode)


def tame_(model, H):
    """SURVEY §8c tamed random init: keeps |z| bounded over 1000 steps of an untrained network so that the timed
    loop runs on finite, representative values (timing itself is data independent)."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("edge_mlp.0.weight") or name.endswith("coord_mlp.0.weight"):
                p[:, 2 * H:] *= 1e-5
        model.dynamics.egnn.embedding_out.weight *= 0.01
        model.dynamics.egnn.embedding_out.bias *= 0.01


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's CPU path) — used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(n_steps, warmup, nodes):
    """Times oracle.sample_p_zs_given_zt (a12) on the host cores; returns seconds per step (median)."""
    O = oracle_module()
    cfg = O.QM9_CFG
    sd = O.make_state_dict(cfg, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    nm, em = O.build_masks(list(nodes), 29)
    bs = len(nodes)
    torch.manual_seed(0)
    noise = O.NoiseSource()
    z = O.combined_noise(cfg, noise, bs, 29, nm, cfg.latent_nf)
    T = cfg.diffusion_steps
    times = []
    with torch.no_grad():
        for k in range(warmup + n_steps):
            s = T - 1 - k
            s_arr = torch.full((bs, 1), float(s)) / T
            t_arr = torch.full((bs, 1), float(s + 1)) / T
            t0 = time.perf_counter()
            z = O.sample_p_zs_given_zt(sd, cfg, s_arr, t_arr, z, nm, em, None, noise)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nodes = workload_nodes(MOLS_PER_GPU * args.gpus)[:CPU_SAMPLE_MOLS]
    steps = max(1, min(args.steps, 8))
    times, cores = cpu_reference_steps(steps, min(args.warmup, 1), nodes)
    sec = float(np.mean(times))
    value = CPU_SAMPLE_MOLS / (FORWARDS_PER_MOLECULE * sec)
    edges = int((nodes * (nodes - 1)).sum())
    sample = (f"{steps} timed sample_p_zs_given_zt steps on the first {CPU_SAMPLE_MOLS} molecules of the workload "
              f"(padded to 29 atoms, fp32, torch CPU ops as in the reference), extrapolated x{FORWARDS_PER_MOLECULE}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "edge_msgs_per_s": 18 * edges / sec,
            "cpu_baseline": {"value": value, "unit": "molecules/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE.json configs[2]: QM9 GeoLDM sampling, egnn_dynamics nf=256 n_layers=9 latent_nf=1, "
                        f"{MOLS_PER_GPU} synthetic molecules/GPU from the QM9 atom-count histogram (<=29 atoms), T=1000, "
                        "random-init weights",
            "molecules_per_gpu": MOLS_PER_GPU, "molecules_total": MOLS_PER_GPU * n_gpus,
            "forwards_per_molecule": FORWARDS_PER_MOLECULE, "parallelism": f"molecule-sharded x{n_gpus}, no collective "
                                                                           "in the loop",
            "l2": "L2 flushed (256 MiB write) between timed steps; step working set ~160 MB > 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)      # ~1 s timed: long enough to sit at the sustained (power-capped) clocks
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mma-mode", default=os.environ.get("GEOLDM_MMA_MODE", "auto"))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mols-per-gpu", type=int, default=MOLS_PER_GPU)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C
    from geoldm_b200 import _lib
    from geoldm_b200.models import get_latent_diffusion
    from geoldm_b200.packing import balance_shards, pack_molecules
    from geoldm_b200.sampling import sample

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world if world > 1 else 1
    W, K = max(args.warmup, 3), max(args.steps, 1)
    mols_per_gpu = args.mols_per_gpu

    L = _lib.lib()
    mode = args.mma_mode
    if mode == "auto":
        mode = "3xf16" if L.geoldm_has_tcgen05() else "fp32"
    margs = qm9_args(mode)
    info = {"atom_decoder": ["H", "C", "N", "O", "F"], "n_nodes": {5: 1}, "max_n_nodes": 29}
    torch.manual_seed(0)
    model, _, _ = get_latent_diffusion(margs, dev, info, None)
    tame_(model, margs.nf)
    model.eval()

    all_nodes = workload_nodes(mols_per_gpu * n_gpus)
    shard = balance_shards(all_nodes, n_gpus)[rank]
    nodes = all_nodes[shard]
    batch = pack_molecules(nodes, dev, mol_ids=shard)
    n_edges = batch.n_edge
    dyn = model.dynamics
    D = 3 + margs.latent_nf
    T = margs.diffusion_steps
    table = model.step_table(dev)
    z = torch.empty(batch.n_node, D, device=dev)
    eps = torch.empty_like(z)
    step_idx = torch.full((1,), T - 1, dtype=torch.int32, device=dev)
    draw_idx = torch.zeros(1, dtype=torch.int32, device=dev)
    cb = batch.c_batch(dyn.egnn.tile_m())
    stream = torch.cuda.Stream(device=dev)

    def sptr():
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def update(mode_):
        _lib.check(L.geoldm_sampler_update(C.byref(cb), mode_, _lib.ptr(table), _lib.ptr(step_idx), _lib.ptr(z),
                                           _lib.ptr(eps), None, 0, D, C.c_uint64(0), _lib.ptr(batch.mol_id),
                                           _lib.ptr(draw_idx), _lib.ptr(z), sptr()), "update")

    def one_step():
        model._denoise_ragged(batch, z, table, step_idx, None, eps)
        update(0)
        _lib.check(L.geoldm_sampler_advance(_lib.ptr(step_idx), -1, _lib.ptr(draw_idx), 1, sptr()), "advance")

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        update(2)
        _lib.check(L.geoldm_sampler_advance(None, 0, _lib.ptr(draw_idx), 1, sptr()), "advance")
        one_step()                                        # lazy init outside capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            one_step()
        clocks = ClockSampler(local)
        clocks.start()
        time.sleep(1.0)                                   # let nvidia-smi come up before anything is timed
        for _ in range(W):                                # untimed warm-up steps
            graph.replay()
        stream.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        n_pre = len(clocks.rows)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for a, b in evs:                                  # EXACTLY K timed steps, L2 flushed between them
            flush.zero_()
            a.record(stream)
            graph.replay()
            b.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        # the timed region is short (K x ~10 ms) against nvidia-smi's 200 ms period: keep the identical load running
        # (untimed graph replays) until at least 6 samples under load exist, then stop the sampler
        t_end = time.time() + 3.0
        while len(clocks.rows) < n_pre + 6 and time.time() < t_end:
            for _ in range(10):
                graph.replay()
            stream.synchronize()
        clocks.rows = clocks.rows[n_pre:]
        clk = clocks.stop()
        clk["note"] = "sampled every 200 ms from the start of the timed steps through an untimed continuation of the same graph replays"
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    finite = bool(torch.isfinite(z).all())

    # ---- dominant kernel (fused GCL edge kernel) timed alone for the roofline ------------------------------------
    H = margs.nf
    w, _keep = dyn.egnn.packed()
    ccfg = dyn.egnn.c_config()
    pq = torch.randn(batch.n_node, 2 * H, device=dev)
    xx = torch.randn(batch.n_node, 3, device=dev)
    agg = torch.zeros(batch.n_node, H, device=dev)
    em = w.block[0].gcl[0].edge
    kev = []
    with torch.cuda.stream(stream):
        for i in range(3 + 10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            _lib.check(L.geoldm_edge_gcl(C.byref(ccfg), C.byref(em), C.byref(cb), _lib.ptr(pq), _lib.ptr(xx),
                                         _lib.ptr(xx), _lib.ptr(agg), sptr()), "edge_gcl")
            b.record(stream)
            if i >= 3:
                kev.append((a, b))
        stream.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    k_flops = n_edges * (2.0 * H * H + 2.0 * H + 4.0 * H)   # second layer + head dot + split first layer adds
    peaks, peak_src = measured_peaks()
    peak_tf = peaks.get("bf16_tflops", 1590.0)              # burst figure: kernel timed alone
    achieved_tf = k_flops / (k_ms * 1e-3) / 1e12
    traffic = None                                          # dram bytes per launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"edge_gcl_{mode}")

    # ---- e2e: one complete sampling job through the public API with host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        nodes_host = torch.from_numpy(nodes.astype(np.int64)).pin_memory()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        one_hot, charges, x, node_mask = sample(margs, dev, model, info, nodesxsample=nodes_host, seed=0, mol_ids=shard)
        out_host = [t.cpu() for t in (one_hot, charges, x)]
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        h2d = nodes_host.numel() * 8
        d2h = sum(t.numel() * t.element_size() for t in out_host)
    # ---- max over ranks -------------------------------------------------------------------------------------------
    stats = torch.tensor([ms_total, e2e_s if not args.no_e2e else 0.0, float(len(nodes)), float(n_edges)],
                         dtype=torch.float64, device=dev)
    if dist is not None:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_max, tot_mols, tot_edges = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    else:
        e2e_max, tot_mols, tot_edges = float(stats[1]), float(stats[2]), float(stats[3])
    ms_per_step = ms_total / K
    value = tot_mols / (FORWARDS_PER_MOLECULE * ms_per_step * 1e-3)
    if not args.no_e2e:
        e2e = {"value": tot_mols / e2e_max, "unit": "molecules/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "what": "one complete sample() job (1000 steps + p(x|z0) + decoder + decode), host nodesxsample in, "
                       "host one_hot/charges/x out; 'step' here = the whole job", "seconds": e2e_max}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            times, cores = cpu_reference_steps(5, 1, all_nodes[:CPU_SAMPLE_MOLS])
            sec = float(np.median(times))
            cpu = {"value": CPU_SAMPLE_MOLS / (FORWARDS_PER_MOLECULE * sec), "unit": "molecules/s", "cores": cores,
                   "kind": "port", "sec_per_step": sec,
                   "sample": f"median of 5 oracle sample_p_zs_given_zt steps on the first {CPU_SAMPLE_MOLS} workload "
                             f"molecules (padded to 29, fp32, torch CPU), extrapolated x{FORWARDS_PER_MOLECULE}"}
        S, Lb = margs.inv_sublayers, margs.n_layers
        if mode == "fp32":      # every GCL / equiv has its own projection launch
            egnn_launches = 1 + Lb * (S * 4 + 3) + 1
        else:                   # equiv P|Q fused with the next block's gcl_0 P|Q
            egnn_launches = 1 + 1 + Lb * (S * 3 + (S - 1) + 3) + 1
            if mode == "3xf16":  # per-edge squared distances: entry coordinates once, current coordinates per later block
                egnn_launches += Lb
        launches_per_step = egnn_launches + 1 + 1 + 2 + 1 + 1   # + prep, nan-flag fill, finish a/b, update, advance
        line = {"metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(n_gpus),
                "mma_mode": mode, "arithmetic": {"fp32": "fp32 FFMA", "3xtf32": "fp32-equivalent products from 3 tf32 MMAs, fp32 accumulate",
                                                  "3xf16": "fp32-equivalent products from 3 fp16 MMAs (hi/lo split operands), fp32 accumulate",
                                                  "tf32": "single tf32 MMA (fails the 1e-5 gate)"}.get(mode, mode),
                "edge_msgs_per_s": 18.0 * tot_edges / (ms_per_step * 1e-3),
                "finite": finite, "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * K,
                "roofline": {"bound": "tensor", "kernel": "fused GCL edge kernel (geoldm_edge_gcl), rank 0 shard",
                             "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                             "traffic": traffic, "kernel_ms": k_ms, "flops_per_launch": k_flops, "peak_source": peak_src},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

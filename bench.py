#!/usr/bin/env python
"""bench.py — GeoLDM QM9 1000-step sampling throughput on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[2], weak scaling): 1250 synthetic QM9 molecules per GPU (10 000 at 8 GPUs),
atom counts drawn from the QM9 histogram (seed 0), EGNN_dynamics nf=256, 9 layers, latent_nf=1, T=1000,
random-init weights (no checkpoints offline).  One bench "step" = one pass of the hot path over the batch =
one ancestral sampling step (denoiser forward + z_s update), replayed from the captured CUDA graph.  A sampled
molecule costs 1002 such passes (1000 steps + p(x|z0) + the same-sized decoder EGNN), so
    value [molecules/s] = molecules_on_all_ranks / (1002 * ms_per_step)          (max over ranks)
`e2e` measures one COMPLETE sampling job (all 1002 passes, decode, host<->device copies) through the public API:
qm9-style `sample(args, device, model, dataset_info, nodesxsample=...)` on one GPU, `distributed.sample_sharded` (same
signature, molecule sharding + the final all_gather) on N > 1.
Extra keys next to the headline (same JSON line, `configs`): the other BASELINE.json configurations measured on the
same box — `latency64` (config 1: 64 molecules, one wave), `geom32` (config 4: GEOM-Drugs shape, batch 32, <= 181
atoms), `train` (config 5: conditional QM9 training step, nf=192, 64 molecules per GPU, NCCL gradient all-reduce).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time
import types

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


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when the
# environment sets NCCL_DEBUG): file descriptor 1 is pointed at stderr for the whole run and the result line goes to the
# saved descriptor.
_RESULT_FD = None


def claim_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def histogram_nodes(hist, n_total, seed=0):
    """Atom counts of synthetic molecules: categorical over an n_nodes histogram (dict order), numpy PCG64."""
    keys = np.array(list(hist.keys()))
    p = np.array(list(hist.values()), dtype=np.float64)
    rng = np.random.default_rng(seed)
    return keys[rng.choice(len(keys), size=n_total, p=p / p.sum())]


def workload_nodes(n_total, seed=0):
    from geoldm_b200.histograms import QM9_WITH_H_N_NODES as hist
    return histogram_nodes(hist, n_total, seed)


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
# Reference arm / CPU baseline: the reference's CPU path of a12 (sample_p_zs_given_zt) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def find_reference():
    """The UNMODIFIED reference tree if this machine has one (build container: /root/reference; a driver-provided
    baseline/_ref otherwise).  It is pure Python without packaging metadata (no setup.py / pyproject.toml), so it cannot
    be pip-installed into baseline/_ref and does not travel to the GPU box: there the oracle port is timed instead."""
    for cand in (os.environ.get("GEOLDM_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "equivariant_diffusion")) and os.path.isdir(os.path.join(cand, "egnn")):
            return cand
    return None


def reference_steps_real(ref_dir, n_steps, warmup, nodes):
    """Times the reference's own EnLatentDiffusion.sample_p_zs_given_zt (en_diffusion.py:716-747) through its public
    constructor qm9/models.py:get_latent_diffusion, default random init, all host threads."""
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):      # plotting imports of qm9/visualizer.py (absent here)
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, ref_dir)
    import contextlib
    import io
    import configs.datasets_config as dc                      # noqa: E402  (reference modules)
    import qm9.models as qm                                   # noqa: E402
    args = qm9_args("fp32")
    args.cuda = False
    del args.mma_mode
    torch.set_num_threads(os.cpu_count() or 1)
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)
        model, _, _ = qm.get_latent_diffusion(args, "cpu", dc.get_dataset_info("qm9", False), None)
    model.eval()
    bs, n_max, T = len(nodes), 29, args.diffusion_steps
    n = torch.as_tensor(np.asarray(nodes)).reshape(-1, 1)
    node_mask = (torch.arange(n_max).unsqueeze(0) < n).float()
    edge_mask = node_mask.unsqueeze(1) * node_mask.unsqueeze(2) * (~torch.eye(n_max, dtype=torch.bool)).unsqueeze(0)
    node_mask, edge_mask = node_mask.unsqueeze(2), edge_mask.reshape(-1, 1)
    times = []
    with torch.no_grad():
        z = model.sample_combined_position_feature_noise(bs, n_max, node_mask)
        for k in range(warmup + n_steps):
            s = T - 1 - (k % T)
            s_arr = torch.full((bs, 1), float(s)) / T
            t_arr = torch.full((bs, 1), float(s + 1)) / T
            t0 = time.perf_counter()
            z = model.sample_p_zs_given_zt(s_arr, t_arr, z, node_mask, edge_mask, None)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def cpu_reference_steps(n_steps, warmup, nodes):
    """Times oracle.sample_p_zs_given_zt (the port of a12) on the host cores; returns seconds per step."""
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
            s = T - 1 - (k % T)
            s_arr = torch.full((bs, 1), float(s)) / T
            t_arr = torch.full((bs, 1), float(s + 1)) / T
            t0 = time.perf_counter()
            z = O.sample_p_zs_given_zt(sd, cfg, s_arr, t_arr, z, nm, em, None, noise)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, same metric, config
    and step / warm-up counts as the GPU arm; each step is a bounded sample of the workload (the first 64 molecules)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nodes = workload_nodes(MOLS_PER_GPU * args.gpus)[:CPU_SAMPLE_MOLS]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = 240.0 / 0.75                                   # ~0.75 s per step on 16 host threads: stay within minutes
    clamped = steps + warm > budget
    if clamped:
        warm = min(warm, 3)
        steps = int(budget) - warm
    ref_dir = find_reference()
    if ref_dir is not None:
        times, cores = reference_steps_real(ref_dir, steps, warm, nodes)
        kind, what = "reference", f"the unmodified reference imported from {ref_dir} (EnLatentDiffusion.sample_p_zs_given_zt)"
    else:
        times, cores = cpu_reference_steps(steps, warm, nodes)
        kind, what = "port", ("oracle/geoldm_oracle.py sample_p_zs_given_zt (torch CPU ops as in the reference; the pure-Python "
                              "reference has no packaging metadata and does not travel to the GPU box)")
    sec = float(np.mean(times))
    value = CPU_SAMPLE_MOLS / (FORWARDS_PER_MOLECULE * sec)
    edges = int((nodes * (nodes - 1)).sum())
    sample = (f"{steps} timed + {warm} warm-up steps of {what} on the first {CPU_SAMPLE_MOLS} molecules of the workload "
              f"(padded to 29 atoms, fp32), extrapolated x{FORWARDS_PER_MOLECULE}"
              + ("; step count clamped to keep the run within minutes" if clamped else ""))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "edge_msgs_per_s": 18 * edges / sec,
            "cpu_baseline": {"value": value, "unit": "molecules/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(n_gpus):
    return {"workload": "BASELINE.json configs[2]: QM9 GeoLDM sampling, egnn_dynamics nf=256 n_layers=9 latent_nf=1, "
                        f"{MOLS_PER_GPU} synthetic molecules/GPU from the QM9 atom-count histogram (<=29 atoms), T=1000, "
                        "random-init weights",
            "molecules_per_gpu": MOLS_PER_GPU, "molecules_total": MOLS_PER_GPU * n_gpus,
            "forwards_per_molecule": FORWARDS_PER_MOLECULE, "parallelism": f"molecule-sharded x{n_gpus}, no collective "
                                                                           "in the loop; final all_gather in e2e",
            "l2": "L2 flushed (256 MiB write) between timed steps; step working set ~160 MB > 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
# timed pieces of the GPU arm
# ------------------------------------------------------------------------------------------------------------------
class StepLoop:
    """One ancestral sampling step over a ragged batch, captured as a CUDA graph (exactly what
    EnLatentDiffusion.sample_latent_ragged replays)."""

    def __init__(self, model, batch, dev, latent_nf, T):
        import ctypes as C
        from geoldm_b200 import _lib
        self.C, self._lib, self.L = C, _lib, _lib.lib()
        self.model, self.batch, self.dev = model, batch, dev
        self.D = 3 + latent_nf
        self.table = model.step_table(dev)
        self.z = torch.empty(batch.n_node, self.D, device=dev)
        self.eps = torch.empty_like(self.z)
        self.step_idx = torch.full((1,), T - 1, dtype=torch.int32, device=dev)
        self.draw_idx = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cb = batch.c_batch(model.dynamics.egnn.tile_m())
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = None

    def _sptr(self):
        return self.C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _update(self, mode_):
        C, _lib = self.C, self._lib
        _lib.check(self.L.geoldm_sampler_update(C.byref(self.cb), mode_, _lib.ptr(self.table), _lib.ptr(self.step_idx),
                                                _lib.ptr(self.z), _lib.ptr(self.eps), None, 0, self.D, C.c_uint64(0),
                                                _lib.ptr(self.batch.mol_id), _lib.ptr(self.draw_idx), _lib.ptr(self.z),
                                                self._sptr()), "update")

    def _one_step(self):
        self.model._denoise_ragged(self.batch, self.z, self.table, self.step_idx, None, self.eps)
        self._update(0)
        self._lib.check(self.L.geoldm_sampler_advance(self._lib.ptr(self.step_idx), -1, self._lib.ptr(self.draw_idx), 1,
                                                      self._sptr()), "advance")

    def capture(self):
        with torch.cuda.stream(self.stream):
            self._update(2)
            self._lib.check(self.L.geoldm_sampler_advance(None, 0, self._lib.ptr(self.draw_idx), 1, self._sptr()), "advance")
            self._one_step()                                  # lazy init outside capture
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._one_step()
        self.stream.synchronize()

    def timed(self, warmup, steps, flush, barrier=None):
        """`steps` graph replays, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.graph.replay()
            self.stream.synchronize()
            if barrier is not None:
                barrier()
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in evs:
                if flush is not None:
                    flush.zero_()
                a.record(self.stream)
                self.graph.replay()
                b.record(self.stream)
            self.stream.synchronize()
            torch.cuda.synchronize()
            if barrier is not None:
                barrier()
        return sum(a.elapsed_time(b) for a, b in evs)


def egnn_launches(mode, n_layers, S):
    """Kernels of THIS repo per denoiser forward (geoldm_egnn_forward, csrc/api.cu) + prep / finish / update / advance."""
    if mode == "fp32":          # every GCL / equiv has its own projection launch
        n = 1 + n_layers * (S * 4 + 3) + 1
    else:                       # equiv P|Q fused with the next block's gcl_0 P|Q
        n = 1 + 1 + n_layers * (S * 3 + (S - 1) + 3) + 1
        if mode == "3xf16":     # per-edge squared distances / normalised differences once per block
            n += n_layers
    return n + 1 + 1 + 2 + 1 + 1


def bench_latency64(margs, model, dev, info):
    """BASELINE.json configs[0]: the headline model on 64 molecules (~158 tiles = one wave): the latency case."""
    from geoldm_b200.packing import pack_molecules
    nodes = workload_nodes(64, seed=1)
    batch = pack_molecules(nodes, dev)
    loop = StepLoop(model, batch, dev, margs.latent_nf, margs.diffusion_steps)
    loop.capture()
    ms = loop.timed(10, 50, None) / 50
    return {"workload": "BASELINE.json configs[0]: QM9 GeoLDM sampling, nf=256 n_layers=9, 64 molecules (one wave)",
            "molecules": 64, "atoms": int(batch.n_node), "edges": int(batch.n_edge), "ms_per_step": ms,
            "molecules_per_s": 64 / (FORWARDS_PER_MOLECULE * ms * 1e-3),
            "edge_msgs_per_s": 18.0 * batch.n_edge / (ms * 1e-3)}


def bench_geom32(mode, dev):
    """BASELINE.json configs[3]: GEOM-Drugs GeoLDM sampling, nf=256 n_layers=4 latent_nf=2, batch 32, <= 181 atoms."""
    from geoldm_b200.histograms import GEOM_WITH_H_N_NODES
    from geoldm_b200.models import get_latent_diffusion
    from geoldm_b200.packing import pack_molecules
    margs = qm9_args(mode)
    margs.n_layers, margs.latent_nf, margs.include_charges, margs.dataset = 4, 2, False, "geom"
    info = {"atom_decoder": list(range(16)), "n_nodes": {44: 1}, "max_n_nodes": 181}
    torch.manual_seed(0)
    model, _, _ = get_latent_diffusion(margs, dev, info, None)
    tame_(model, margs.nf)
    model.eval()
    nodes = histogram_nodes(GEOM_WITH_H_N_NODES, 32, seed=0)
    nodes[0] = 181                                             # the batch contains the largest molecule of the dataset
    batch = pack_molecules(nodes, dev)
    loop = StepLoop(model, batch, dev, margs.latent_nf, margs.diffusion_steps)
    loop.capture()
    ms = loop.timed(10, 50, None) / 50
    return {"workload": "BASELINE.json configs[3]: GEOM-Drugs GeoLDM sampling, nf=256 n_layers=4 latent_nf=2, batch 32 "
                        "from the GEOM atom-count histogram (incl. one 181-atom molecule)",
            "molecules": 32, "atoms": int(batch.n_node), "edges": int(batch.n_edge), "max_atoms": int(nodes.max()),
            "ms_per_step": ms, "molecules_per_s": 32 / (FORWARDS_PER_MOLECULE * ms * 1e-3),
            "edge_msgs_per_s": 8.0 * batch.n_edge / (ms * 1e-3)}


def bench_train(mode, dev, rank, world, dist, steps=6):
    """BASELINE.json configs[4]: conditional QM9 training step (nf=192, 9 layers, alpha conditioning, 64 molecules per
    GPU, trainable first stage) incl. the NCCL gradient all-reduce; ms/step = max over ranks."""
    from geoldm_b200 import losses, training
    from geoldm_b200.histograms import HISTOGRAMS
    from geoldm_b200.models import get_latent_diffusion
    from geoldm_b200.sampling import build_masks
    bs = 64
    args = qm9_args(mode)
    args.include_charges, args.context_node_nf, args.nf, args.normalize_factors = False, 1, 192, [1, 8, 1]
    args.trainable_ae, args.dataset, args.lr, args.clip_grad, args.ode_regularization = True, "qm9_second_half", 1e-4, True, 0.0
    hist = HISTOGRAMS["qm9_second_half"]
    info = {"atom_decoder": list(range(5)), "n_nodes": hist, "max_n_nodes": 29}
    torch.manual_seed(0)                                        # identical init on every rank
    model, nodes_dist, _ = get_latent_diffusion(args, dev, info, None)
    # the global batch of world x 64 molecules, dealt over the ranks with equal molecule counts and balanced edge counts
    # (packing.balance_shards_equal: the step is as slow as its slowest rank); one rank: the batch as drawn
    from geoldm_b200.packing import balance_shards_equal
    all_nodes = np.asarray(histogram_nodes(hist, world * bs, seed=5))
    nodes = all_nodes[balance_shards_equal(all_nodes, world)[rank]] if world > 1 else all_nodes
    gen = torch.Generator().manual_seed(11 + rank)
    nm, em = build_masks(torch.as_tensor(nodes), 29, dev)
    x = losses.remove_mean_with_mask(torch.randn(bs, 29, 3, generator=gen).to(dev) * nm, nm)
    one_hot = torch.nn.functional.one_hot(torch.randint(0, 5, (bs, 29), generator=gen).to(dev), 5).float() * nm
    ctx = torch.randn(bs, 1, 1, generator=gen).to(dev).expand(-1, 29, -1) * nm
    h = {"categorical": one_hot, "integer": torch.zeros(0, device=dev)}
    x_host = x.cpu().pin_memory()
    optim = training.get_optim(args, model, capturable=True)
    model_ema = copy.deepcopy(model)
    ema = training.EMA(args.ema_decay)
    buckets = training.FlatGradBuckets(model)                  # flat per-bucket gradients, all-reduce overlapped with backward
    nbytes = buckets.nbytes
    # the whole step (loss, backward, bucketed all-reduce, device-side adaptive clipping, AdamW, EMA) as ONE captured graph
    step = training.GraphedTrainStep(args, model, optim, nodes_dist, x, h, nm, em, ctx, model_ema=model_ema, ema=ema,
                                     buckets=buckets)
    times = []
    for it in range(steps + 2):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nll, gn = step(x_host)                                 # the step's positions arrive from pinned host memory
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = torch.tensor([float(np.mean(times))], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    last_nll = float(nll)
    step.close()                                               # the graph holds NCCL work: release it before the process group goes
    nll = torch.tensor(last_nll)
    return {"workload": "BASELINE.json configs[4]: conditional QM9 GeoLDM training step, nf=192 n_layers=9, 64 molecules "
                        "per GPU (the global batch dealt over the ranks by edge count, equal molecule counts), trainable first stage, "
                        "AdamW + EMA, gradient all-reduce over NCCL; the step is one captured CUDA graph (training.GraphedTrainStep)",
            "molecules_per_gpu": bs, "n_gpus": world, "ms_per_step": ms, "molecules_per_s": world * bs / (ms * 1e-3),
            "allreduce_bytes_per_step": nbytes if world > 1 else 0, "last_nll": float(nll), "finite": bool(torch.isfinite(nll))}


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
    ap.add_argument("--no-extra", action="store_true", help="skip the configs.latency64 / geom32 / train measurements")
    ap.add_argument("--mols-per-gpu", type=int, default=MOLS_PER_GPU)
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C
    from geoldm_b200 import _lib
    from geoldm_b200.distributed import sample_sharded
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
    loop = StepLoop(model, batch, dev, margs.latent_nf, margs.diffusion_steps)
    loop.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    barrier = dist.barrier if dist is not None else None

    # ---- headline: EXACTLY K timed steps after W warm-up steps, clocks sampled during the timed region ---------------
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(1.0)                                       # let nvidia-smi come up before anything is timed
    n_pre = [None]

    def barrier_and_mark():
        if barrier is not None:
            barrier()
        if n_pre[0] is None:                              # first call = start of the timed steps
            n_pre[0] = len(clocks.rows)

    ms_total = loop.timed(W, K, flush, barrier_and_mark)
    # the timed region is short (K x ~5 ms) against nvidia-smi's 200 ms period: keep the identical load running
    # (untimed graph replays) until at least 6 samples under load exist, then stop the sampler
    t_end = time.time() + 3.0
    with torch.cuda.stream(loop.stream):
        while len(clocks.rows) < n_pre[0] + 6 and time.time() < t_end:
            for _ in range(10):
                loop.graph.replay()
            loop.stream.synchronize()
    clocks.rows = clocks.rows[n_pre[0]:]
    clk = clocks.stop()
    clk["note"] = "sampled every 200 ms from the start of the timed steps through an untimed continuation of the same graph replays"
    finite = bool(torch.isfinite(loop.z).all())

    # ---- dominant kernel (fused GCL edge kernel) timed alone for the roofline, through the entry point and with the
    # inputs (precomputed per-edge distances) that geoldm_egnn_forward uses ------------------------------------------
    H = margs.nf
    w, _keep = dyn.egnn.packed()
    ccfg = dyn.egnn.c_config()
    cb = loop.cb
    pq = torch.randn(batch.n_node, 2 * H, device=dev)
    xx = torch.randn(batch.n_node, 3, device=dev)
    agg = torch.zeros(batch.n_node, H, device=dev)
    em = w.block[0].gcl[0].edge
    stream = loop.stream
    pre = mode == "3xf16"
    kev = []
    with torch.cuda.stream(stream):
        sp = C.c_void_p(stream.cuda_stream)
        if pre:
            r_e = torch.empty(n_edges, device=dev)
            _lib.check(L.geoldm_edge_dist(C.byref(cb), _lib.ptr(xx), _lib.ptr(r_e), None, 1.0, sp), "edge_dist")
        for i in range(3 + 10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            if pre:
                _lib.check(L.geoldm_edge_gcl_pre(C.byref(ccfg), C.byref(em), C.byref(cb), _lib.ptr(pq), 2 * H, _lib.ptr(r_e),
                                                 _lib.ptr(r_e), _lib.ptr(agg), sp), "edge_gcl_pre")
            else:
                _lib.check(L.geoldm_edge_gcl(C.byref(ccfg), C.byref(em), C.byref(cb), _lib.ptr(pq), _lib.ptr(xx),
                                             _lib.ptr(xx), _lib.ptr(agg), sp), "edge_gcl")
            b.record(stream)
            if i >= 3:
                kev.append((a, b))
        stream.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    k_flops = n_edges * (2.0 * H * H + 2.0 * H)              # second edge layer + attention head dot (tensor-shaped work only)
    peaks, peak_src = measured_peaks()
    peak_tf = peaks.get("bf16_tflops", 1590.0)              # burst figure: kernel timed alone
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved_tf = k_flops / (k_ms * 1e-3) / 1e12
    traffic = None                                          # dram bytes per launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"edge_gcl_{mode}")

    # ---- e2e: one complete sampling job through the public API with host buffers ---------------------------------
    e2e = None
    e2e_s = 0.0
    if not args.no_e2e:
        # one small untimed job first (the e2e counterpart of the warm-up steps): lazy one-time work - decoder weight
        # images, kernel attributes, allocator pools - is not part of a job's steady-state cost; the timed job still
        # packs its own batch, captures its own CUDA graph and copies its inputs / outputs
        wn = torch.from_numpy(all_nodes[:8 * world].astype(np.int64) if dist is not None else nodes[:8].astype(np.int64)).pin_memory()
        if dist is None:
            sample(margs, dev, model, info, nodesxsample=wn, seed=1)
        else:
            sample_sharded(margs, dev, model, info, wn, seed=1)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        if dist is None:
            nodes_host = torch.from_numpy(nodes.astype(np.int64)).pin_memory()
            one_hot, charges, x, node_mask = sample(margs, dev, model, info, nodesxsample=nodes_host, seed=0, mol_ids=shard)
            n_in = nodes_host.numel()
        else:                   # N ranks: every rank passes the WHOLE job; sharding + the final all_gather are inside
            nodes_host = torch.from_numpy(all_nodes.astype(np.int64)).pin_memory()
            one_hot, charges, x, node_mask = sample_sharded(margs, dev, model, info, nodes_host, seed=0)
            n_in = len(nodes)
        out_host = [t.cpu() for t in (one_hot, charges, x)]
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        h2d = n_in * 8
        d2h = sum(t.numel() * t.element_size() for t in out_host)
        gathered = 0 if dist is None else sum(t.numel() * t.element_size() for t in (one_hot, charges, x, node_mask))

    # ---- the other BASELINE configurations, on the same box (extra keys; the headline stays config 3) ------------------
    extra = {}
    if not args.no_extra:
        if rank == 0:
            try:
                extra["latency64"] = bench_latency64(margs, model, dev, info)
            except Exception as e:          # an extra line must never take the headline down
                extra["latency64"] = {"error": repr(e)}
            try:
                extra["geom32"] = bench_geom32(mode, dev)
            except Exception as e:
                extra["geom32"] = {"error": repr(e)}
        try:
            extra["train"] = bench_train(mode, dev, rank, world, dist)
        except Exception as e:
            if dist is not None:
                raise
            extra["train"] = {"error": repr(e)}

    # ---- max over ranks -------------------------------------------------------------------------------------------
    stats = torch.tensor([ms_total, e2e_s, float(len(nodes)), float(n_edges)], dtype=torch.float64, device=dev)
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
               "what": "one complete sampling job (1000 steps + p(x|z0) + decoder + decode), host nodesxsample in, host "
                       "one_hot/charges/x out; 'step' here = the whole job; N > 1: distributed.sample_sharded, i.e. the "
                       "final all_gather of every rank's molecules is inside the timed region",
               "seconds": e2e_max, "all_gather_bytes_per_rank": gathered,
               "warmup": "one untimed 8-molecule-per-rank job before the timed one (lazy one-time initialisation)"}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            times, cores = cpu_reference_steps(5, 1, all_nodes[:CPU_SAMPLE_MOLS])
            sec = float(np.median(times))
            cpu = {"value": CPU_SAMPLE_MOLS / (FORWARDS_PER_MOLECULE * sec), "unit": "molecules/s", "cores": cores,
                   "kind": "port", "sec_per_step": sec,
                   "sample": f"median of 5 oracle sample_p_zs_given_zt steps on the first {CPU_SAMPLE_MOLS} workload "
                             f"molecules (padded to 29, fp32, torch CPU), extrapolated x{FORWARDS_PER_MOLECULE}"}
        launches_per_step = egnn_launches(mode, margs.n_layers, margs.inv_sublayers)
        line = {"metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(n_gpus),
                "mma_mode": mode, "arithmetic": {"fp32": "fp32 FFMA", "3xtf32": "fp32-equivalent products from 3 tf32 MMAs, fp32 accumulate",
                                                  "3xf16": "fp32-equivalent products from 3 fp16 MMAs (hi/lo split operands), fp32 accumulate",
                                                  "tf32": "single tf32 MMA (fails the 1e-5 gate)"}.get(mode, mode),
                "edge_msgs_per_s": 18.0 * tot_edges / (ms_per_step * 1e-3),
                "finite": finite, "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * K,
                "roofline": {"bound": "tensor", "kernel": "fused GCL edge kernel (geoldm_edge_gcl_pre: the launch of "
                                                           "geoldm_egnn_forward, precomputed distances), rank 0 shard",
                             "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                             "frac_of_sustained_peak": achieved_tf / peak_sus, "peak_sustained": peak_sus,
                             "traffic": traffic, "kernel_ms": k_ms, "flops_per_launch": k_flops,
                             "flops": "E (2 H^2 + 2 H): second edge layer + head dot; 3 fp16 MMAs are issued per product",
                             "peak_source": peak_src},
                "cpu_baseline": cpu, "configs": extra}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

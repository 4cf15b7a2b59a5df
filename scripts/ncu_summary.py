"""Condense an ncu report (--set full) into the per-launch summary committed under profiles/:
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_kernels_ncu_summary.csv
One row per profiled launch: duration, tensor / XU / FMA / ALU pipe activity, issue slots, L1 data-pipe wavefronts, DRAM bytes,
registers, and the warp-stall sample shares (long scoreboard, wait, short scoreboard, mio throttle, barrier)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("kernel", "Kernel Name"), ("duration_us", "gpu__time_duration.sum"),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue_slots_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
        ("l1_lsu_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("warp_inst_executed", "smsp__inst_executed.sum"),
        ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("registers", "launch__registers_per_thread"), ("threads", "launch__block_size"),
        ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic")]
stalls = ["long_scoreboard", "wait", "short_scoreboard", "mio_throttle", "barrier", "math_pipe_throttle", "not_selected",
          "selected", "sleeping", "branch_resolving"]
out = csv.writer(sys.stdout)
out.writerow([w[0] for w in want] + ["stall_" + s + "_pct" for s in stalls])
for r in rows[2:]:
    vals = []
    for name, key in want:
        v = r[col[key]] if key in col else ""
        if name == "kernel":
            v = v.replace("void geoldm::<unnamed>::", "").replace("(geoldm::<unnamed>::Args)", "")
        vals.append(v)
    tot = 0.0
    sv = []
    for h, i in col.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
            tot += float(r[i] or 0)
    for s_ in stalls:
        k = "smsp__pcsamp_warps_issue_stalled_" + s_
        sv.append(f"{100 * float(r[col[k]] or 0) / tot:.1f}" if k in col and tot else "")
    out.writerow(vals + sv)

#!/usr/bin/env python
"""One ncu capture of a bench workload's dominant kernel -> profiles/r02_ncu_<workload>.json, the file bench.py reads
its `roofline.traffic` and hardware FP64 counts from.  The JSON records the hash of the CUDA sources the library was
built from (bench.source_hash): bench.py refuses a capture taken from other sources.

  python scripts/ncu_profile.py c3            # on the GPU box (under gpurun); writes profiles/ and gpurun_out/
  python scripts/ncu_profile.py c5 --iters 10 # long kernels: capture fewer iterations (bench.py scales by chain-steps)

Numbers under ncu are never bench values: only counters (bytes, instructions, pipe utilisation) are kept; the duration
is recorded for the kernel's share, not for throughput.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__inst_executed.sum",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def child(name, chains, iters):
    import mcmcf90_b200 as mb
    W = bench.Workload(name)
    s = mb.Sampler(mb.default_config(nchains=chains, seed=bench.SEED, nsimu=1 + iters * 8, model=W.model, pool_adapt=W.pool, **W.nml))
    s.set_data(W.blob(mb.models))
    s.set_initial(W.par0(chains, 0), W.cmat0, W.sigma2, W.nobs)
    s.run(plan(W, iters)[0])  # one launch of the step kernel, then (warp-per-chain kernels) the first adaptation tick
    s.run(plan(W, iters)[1])  # <- the captured launch: the step kernel of this call, no tick inside
    s.close()


def plan(W, iters):
    """(warm-up iterations, captured iterations): the warp-per-chain kernels are launched per segment between adaptation
    ticks, so the warm-up stops right at the first tick and the captured launch stays inside the next interval."""
    if W.kernel == "k1_step_kernel":
        return max(iters, W.nml.get("adaptint", iters)), iters
    a = W.nml["adaptint"]
    if W.kernel == "k4_ram_step_kernel":  # one launch per mcmcb_run piece; pooled runs are cut at the pooled ticks
        return a - 1, min(iters, a)
    return a - 1, min(iters, a - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    W = bench.Workload(a.workload)
    chains, iters = a.chains or W.chains, a.iters or W.iters
    if a.child:
        return child(a.workload, chains, iters)
    warm, iters = plan(W, iters)
    cmd = ["ncu", "--csv", "--print-units", "base", "--metrics", ",".join(METRICS), "--clock-control", "none",
           "-k", "regex:" + W.kernel, "-s", "1", "-c", "1", sys.executable, os.path.abspath(__file__), a.workload, "--child",
           "--chains", str(chains), "--iters", str(iters)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    text = r.stdout
    start = text.find('"ID"')
    if r.returncode != 0 or start < 0:
        sys.stderr.write(text[-3000:] + r.stderr[-3000:])
        return 1
    vals, kname = {}, None
    for row in csv.DictReader(io.StringIO(text[start:])):
        kname = row["Kernel Name"]
        vals[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    g = lambda k: vals.get(k, 0.0)  # noqa: E731
    out = {
        "workload": a.workload, "kernel": kname, "chains": chains, "iterations": iters, "warmup_iterations": warm,
        "source_hash": bench.source_hash(),
        "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
        "dram_bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
        "duration_ms": g("gpu__time_duration.sum") / 1e6,
        "inst_dfma": g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"),
        "inst_dadd": g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"),
        "inst_dmul": g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
        "fp64_flops": 2 * g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum") +
                      g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum") + g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
        "warp_inst_executed": g("smsp__inst_executed.sum"),
        "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "shared_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "shared_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "l2_hit_pct": g("lts__t_sector_hit_rate.pct"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "command": " ".join(cmd[:cmd.index(sys.executable)]) + " python scripts/ncu_profile.py %s --child ..." % a.workload,
        "note": "cold-cache, serialised replay passes: counters only; never a throughput number",
    }
    for d in ("profiles", "gpurun_out"):
        os.makedirs(os.path.join(ROOT, d), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, d, "r02_ncu_%s.json" % a.workload), "w"), indent=1)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())

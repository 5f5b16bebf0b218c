#!/usr/bin/env python
"""bench.py -- chain-steps/s of the batched adaptive-MH hot path on the BASELINE config C3.

Workload (BASELINE.json configs[2], the config the whole-box target is quoted on; SURVEY.md
8d): 2^20 independent DRAM+AM chains PER GPU (weak scaling) on the exponential-regression
model y = th1*exp(-th2*x) with ndata = 10^4 observations staged in shared memory,
drscale=2, adaptint=100, initcmatn=1, sigma2 Gibbs update on.  One bench "step" = one pass
of the hot path = ONE kernel launch advancing every chain by 100 MCMC iterations
(one adaptation interval).  Data are synthetic (seeded); FP64 throughout.

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference

`value`  : device-timed (CUDA events on the kernel's stream, max over ranks), state resident in HBM.
`e2e`    : same metric through the C ABI with HOST buffers: per step H2D of every chain's start
           point (pinned), the kernel, D2H of theta/mean/cov/counters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NDATA = 10000
MCMC_PER_STEP = 100
NML = dict(adaptint=100, drscale=2.0, initcmatn=1, doburnin=0, burnintime=0, updatesigma=1, N0=1.0, S02=0.5)
SEED = 2024
PAR0 = np.array([10.0, 0.1])
CMAT0 = np.diag([0.2, 0.001]) * (11.0 / NDATA)


def synth_data():
    rng = np.random.default_rng(SEED)
    x = 10.0 * np.arange(NDATA) / (NDATA - 1)
    y = 10.0 * np.exp(-0.1 * x) + rng.normal(0.0, np.sqrt(0.5), NDATA)
    return x, y


def start_points(n, offset):
    """par0 = (10, 0.1) with 1% jitter per chain, keyed by the global chain id block."""
    rng = np.random.default_rng([SEED, offset])
    return PAR0 * (1.0 + 0.01 * rng.standard_normal((n, 2)))


def flops_per_chain_step(q, rows_per_step, d=2, n=NDATA, adaptint=100):
    """Algorithmic FP64 flops of one DRAM/AM chain-step (exp counted as 1)."""
    f_prop, f_ss, f_q1, f_alpha = d * (d + 1), 6 * n, 4 * d * d + 6 * d, 12
    f_adapt = 5 * d * d * rows_per_step * adaptint + d ** 3 / 3 + 2 * d ** 3 / 3
    return (1 + q) * (f_prop + f_ss + 3 * d) + q * f_q1 + (1 + q) * f_alpha + f_adapt / adaptint


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------- reference arm
def cpu_baseline(cores, target_seconds=15.0, chains_per_core=4):
    """The oracle (C restatement of the reference, 'port') timed on the host cores: one chain
    per thread at a time (the reference is one chain per process)."""
    from oracle import oracle as O
    x, y = synth_data()
    blob = O.blob_expreg(x, y)
    nch = cores * chains_per_core

    def run(nsimu):
        cfg = O.make_cfg(nsimu=nsimu, **NML)
        out = O.run_batch(cfg, O.MODEL_EXPREG, blob, start_points(nch, 0), CMAT0, [0.5], [NDATA], seed=SEED,
                          chain0=0, nthreads=cores)
        return out["seconds"]

    t = run(51)  # calibration
    rate = nch * 50 / max(t, 1e-6)
    nsimu = int(max(101, min(200000, target_seconds * rate / nch))) + 1
    sec = run(nsimu)
    return nch * (nsimu - 1) / sec, "%d chains x %d MCMC steps of C3 (ndata=%d), %d threads, %.1f s" % (
        nch, nsimu - 1, NDATA, cores, sec)


def run_reference(args):
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return 0
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    x, y = synth_data()
    blob = O.blob_expreg(x, y)
    nch, msteps = 8 * cores, 500  # bounded sample of one C3 step per bench step
    cfg = O.make_cfg(nsimu=msteps + 1, **NML)
    par0 = start_points(nch, 0)
    times = []
    for it in range(args.warmup + args.steps):
        out = O.run_batch(cfg, O.MODEL_EXPREG, blob, par0, CMAT0, [0.5], [NDATA], seed=SEED + it, chain0=0,
                          nthreads=cores)
        if it >= args.warmup:
            times.append(out["seconds"])
    total = sum(times)
    value = nch * msteps * args.steps / total
    sample = "%d chains x %d MCMC steps per step on %d host threads (bounded sample of the 2^20-chain step)" % (
        nch, msteps, cores)
    line = {
        "impl": "reference", "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference (oracle/): the Fortran reference cannot be built here (no Fortran compiler)",
    }
    print(json.dumps(line))
    return 0


def other_workloads(mb, torch, dev):
    """Short device-timed runs of the other BASELINE configurations' shapes (C2, C4, C5: the large-npar kernels) on this
    GPU.  Not part of the headline metric: reported beside it so that every configuration has a measured number in
    the bench record.  Algorithmic HBM bytes per step as in SURVEY.md 8d (private factor per chain)."""
    def gauss_target(d, rho=0.9):
        sd = 1.0 + 9.0 * np.arange(d) / max(d - 1, 1)
        sig = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d))) * np.outer(sd, sd)
        lam = np.linalg.inv(sig)
        return np.zeros(d), 0.5 * (lam + lam.T)

    def timeit(label, cfg_kw, model, blob, d, n, steps, cmat0, bytes_per_step):
        s = mb.Sampler(mb.default_config(nchains=n, seed=12345, model=model, device=dev, **cfg_kw))
        s.set_data(blob)
        s.set_initial(np.zeros(d), cmat0, [1.0], [1])
        s.run(steps)  # warm-up incl. the initial evaluation
        st = torch.cuda.ExternalStream(s.stream, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            c0 = s.counters()
            e0.record(st)
            s.run(steps, sync=False)
            e1.record(st)
            s.sync()
            ms = e0.elapsed_time(e1)
            c1 = s.counters()
            if best is None or ms < best[0]:
                best = (ms, float((c1["drtries"] - c0["drtries"]).sum()) / (n * steps), int((c1["status"] != 0).sum()))
        info = s.info()
        s.close()
        ms, q, bad = best
        rate = n * steps / ms * 1e3
        return {"workload": label, "npar": d, "chains": n, "iterations_timed": steps, "ms": ms, "chain_steps_per_s": rate,
                "stage2_rate_q": q, "algorithmic_hbm_bytes_per_step": bytes_per_step(q),
                "algorithmic_GBps": rate * bytes_per_step(q) / 1e9, "threads_per_chain": info["lanes_per_chain"],
                "chains_with_error_status": bad}

    out = []
    # C1 batched: the reference's own testcase (11 data) run as 2^22 chains -- the small-ndata end of the metric
    x11 = np.arange(11.0)
    y11 = np.array([9.33, 9.40, 8.99, 7.06, 7.13, 6.69, 4.69, 4.24, 4.77, 3.86, 4.02])
    n1 = 1 << 22
    s = mb.Sampler(mb.default_config(nchains=n1, seed=3, model="expreg", device=dev, nsimu=100000, adaptint=100, drscale=2.0,
                                     initcmatn=1, updatesigma=1, N0=1.0, S02=0.0))
    s.set_data(mb.models.blob_expreg(x11, y11))
    s.set_initial(np.array([10.0, 0.1]), np.diag([0.2, 0.001]), [0.5], [11])
    s.run(200)
    st = torch.cuda.ExternalStream(s.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = s.counters()
    e0.record(st)
    s.run(200, sync=False)
    e1.record(st)
    s.sync()
    c1 = s.counters()
    ms = e0.elapsed_time(e1)
    q1 = float((c1["drtries"] - c0["drtries"]).sum()) / (n1 * 200)
    out.append({"workload": "C1 batched: testcases/data.dat model (11 data), DRAM+AM, 2^22 chains (k1_step_kernel, one chain per "
                            "thread)", "npar": 2, "chains": n1, "iterations_timed": 200, "ms": ms,
                "chain_steps_per_s": n1 * 200 / ms * 1e3, "stage2_rate_q": q1,
                "datum_evals_per_s": n1 * 200 * (1 + q1) * 11 / ms * 1e3, "chains_per_thread": s.info()["chains_per_thread"],
                "chains_with_error_status": int((c1["status"] != 0).sum())})
    s.close()
    d = 100
    mu, lam = gauss_target(d)
    tri = d * (d + 1) // 2
    out.append(timeit("C2: 4096 DRAM chains, 100-dim correlated Gaussian (k2_step_kernel + k2_adapt_kernel at the tick)",
                      dict(nsimu=100000, adaptint=200, drscale=2.0, initcmatn=1, updatesigma=0), "gauss",
                      mb.models.blob_gauss(mu, lam), d, 4096, 200, 0.01 * np.eye(d),
                      lambda q: 8 * (tri * (1 + q) + 3 * d + 10)))
    d = 50
    tri = d * (d + 1) // 2
    out.append(timeit("C4: 65536 RAM chains, 50-dim banana (k2_step_kernel, per-chain factor; pooling off)",
                      dict(method=mb.RAM, nsimu=100000, updatesigma=0, alphatarget=0.234, nuparam=0.7), "banana",
                      mb.models.blob_banana(d, 0.03), d, 65536, 100, np.eye(d), lambda q: 16 * tri + 8 * (3 * d + 10)))
    groups, per = 198, 10
    d = groups + 2
    rng = np.random.default_rng(5)
    y = rng.normal(size=(groups, 1)) + rng.normal(size=(groups, per))
    out.append(timeit("C5 shape at 2048 chains: SCAM, 200-param hierarchical model (k3_scam_step_kernel; one step = a sweep "
                      "over the 200 components; per-chain 320 KB rotation)",
                      dict(method=mb.SCAM, nsimu=100000, adaptint=100, initcmatn=1, updatesigma=0), "hier",
                      mb.models.blob_hier(y), d, 2048, 20, 0.1 * np.eye(d), lambda q: 8 * (d * d + 3 * d + 10)))
    return out


def workload_config(args):
    return {"workload": "C3: DRAM+AM chains on exp-regression, ndata=%d in shared memory" % NDATA,
            "chains_per_gpu": args.chains, "mcmc_iterations_per_step": MCMC_PER_STEP, "npar": 2,
            "namelist": NML, "rng": "philox4x32-10", "parallelism": "chains sharded over GPUs, no collective",
            "l2": "per-chain state (~220 MB/GPU) exceeds L2; no flush needed"}


# --------------------------------------------------------------------------- this framework
def run_ours(args):
    import torch
    import mcmcf90_b200 as mb

    rank, world, local = dist_setup(args.gpus)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)
    N = args.chains
    offset = rank * N
    x, y = synth_data()
    blob = mb.models.blob_expreg(x, y)
    nsimu = 1 + MCMC_PER_STEP * (2 * (args.warmup + args.steps) + 8)
    cfg = mb.default_config(nchains=N, chain_offset=offset, seed=SEED, device=dev, nsimu=nsimu, model="expreg",
                            lanes_per_chain=args.lanes, **NML)
    s = mb.Sampler(cfg)
    s.set_data(blob)
    par0_pinned = torch.empty((N, 2), dtype=torch.float64).pin_memory()
    par0 = par0_pinned.numpy()
    par0[:] = start_points(N, offset)
    s.set_initial(par0, CMAT0, [0.5], [NDATA])
    stream = torch.cuda.ExternalStream(s.stream, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # FP64 pipe peak, measured on this GPU: burst (best 22 ms run) and sustained (2 s of back-to-back runs)
    burst = mb.dfma_peak(dev)[0]
    t_end, sus = time.time() + 2.0, []
    while time.time() < t_end:
        sus.append(mb.dfma_peak(dev)[0])
    sustained = float(np.mean(sus[len(sus) // 2:]))

    for _ in range(args.warmup):
        s.run(MCMC_PER_STEP, sync=False)
    s.sync()
    c0 = s.counters()
    l0 = s.launches
    clocks = ClockSampler(dev)
    barrier()
    clocks.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record(stream)
    for a, b in evs:
        a.record(stream)
        s.run(MCMC_PER_STEP, sync=False)
        b.record(stream)
    e_all1.record(stream)
    s.sync()
    barrier()
    clk = clocks.stop()
    total_ms = e_all0.elapsed_time(e_all1)
    launch_ms = [a.elapsed_time(b) for a, b in evs]
    launches = s.launches - l0
    c1 = s.counters()
    done_steps = MCMC_PER_STEP * args.steps
    q = float((c1["drtries"] - c0["drtries"]).sum()) / (N * done_steps)
    rows = float((c1["chainind"] - c0["chainind"]).sum()) / (N * done_steps)
    stay = float((c1["stayed"] - c0["stayed"]).sum()) / (N * done_steps)
    status_bad = int((c1["status"] != 0).sum())

    # ---- end to end through the C ABI with host buffers
    h2d = par0.nbytes + blob.nbytes + CMAT0.nbytes + 8 + 4
    d2h = 0
    e2e_sampler = mb.Sampler(mb.default_config(nchains=N, chain_offset=offset, seed=SEED + 1, device=dev,
                                               nsimu=MCMC_PER_STEP + 1, model="expreg", lanes_per_chain=args.lanes,
                                               **NML))
    e2e_sampler.set_data(blob)
    # results land in pinned host buffers (caller-owned, as the C ABI prescribes)
    outs = {"par": torch.empty((N, 2), dtype=torch.float64).pin_memory(),
            "mean": torch.empty((N, 2), dtype=torch.float64).pin_memory(),
            "cmat": torch.empty((N, 4), dtype=torch.float64).pin_memory(),
            "counters": torch.empty((N, 8), dtype=torch.int64).pin_memory()}
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for it in range(e2e_steps):
        e2e_sampler.set_data(blob)                                   # H2D model data
        e2e_sampler.set_initial(par0, CMAT0, [0.5], [NDATA])         # H2D start points (pinned) + init kernel
        e2e_sampler.run(MCMC_PER_STEP, sync=False)
        out = [e2e_sampler.fetch(w, out=t.numpy()) for w, t in outs.items()]  # D2H results
        d2h = sum(o.nbytes for o in out)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_sampler.close()

    t_total = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda:%d" % dev)
    if dist is not None:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = [float(v) for v in t_total.cpu()]
    value = world * N * done_steps / (total_ms_max * 1e-3)
    e2e_value = world * N * MCMC_PER_STEP * e2e_steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        fl = flops_per_chain_step(q, rows)
        avg_launch_ms = float(np.mean(launch_ms))
        achieved = N * MCMC_PER_STEP * fl / (avg_launch_ms * 1e-3) / 1e12
        info = s.info()
        # hardware FP64 instruction count per datum of the compiled ssfunction loop (see DESIGN.md;
        # DFMA counted as 2 flops) -- explains the gap between algorithmic and pipe utilisation
        # 9 FP64 instructions per datum in the SASS of the datum loop (7 DFMA + 1 DMUL + 1 DADD, profiles/r01_summary.md G)
        hw_flop_per_datum = float(os.environ.get("MCMCB_HW_FLOP_PER_DATUM", "16") or 0)
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE steady-state bench launch (2^20 chains x 100 iterations),
        # from `ncu --metrics dram__bytes...` on this command (profiles/r01_bench_launch_dram.txt).  The algorithmic
        # state traffic is 2 x 212 B x 2^20 = 0.44 GB; the rest is the chains' cold state (accept/reject, adaptation,
        # RNG position: ~450 B per chain, in local memory) cycling through L2: with 4 chains per thread the 3 x 10^5
        # chains in flight hold 136 MB of it, more than L2 keeps, so every step re-reads and re-writes it.  That is
        # 46 GB/s = 0.7 % of the measured HBM bandwidth on a kernel bound by the FP64 pipe and shared memory
        # (one chain per thread: 0.74 GB per launch, 7 % slower).
        traffic = {4: 58.5e9, 1: 738.7e6}.get(info["chains_per_thread"]) if (N == 1 << 20 and info["lanes_per_chain"] == 1) else None
        roof = {"bound": "fp64", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": achieved / sustained, "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu, profiles/r01_bench_launch_dram.txt)",
                "algorithmic_bytes_per_launch": 2.0 * 212 * N,
                "peak_source": "in-bench DFMA microbenchmark on this GPU, sustained 2 s (burst %.2f); "
                               "MEASURED_PEAKS.json holds only HBM and bf16-tensor peaks, neither bounds this kernel" % burst,
                "algorithmic_flops_per_chain_step": fl, "stage2_rate_q": q, "accept_rate": 1 - stay,
                "datum_evals_per_s": N * MCMC_PER_STEP * (1 + q) * NDATA / (avg_launch_ms * 1e-3),
                "kernel": "k1_step_kernel<ExpReg,L=%d,smem,B=%d>" % (info["lanes_per_chain"], info["chains_per_thread"]),
                "avg_launch_ms": avg_launch_ms, "launch_ms": launch_ms}
        if hw_flop_per_datum > 0:
            hw = roof["datum_evals_per_s"] * hw_flop_per_datum / 1e12
            roof["achieved_hw"] = hw
            roof["frac_hw"] = hw / sustained
            roof["hw_note"] = ("FP64 flops the compiled datum loop executes (exp = 7 FP64 instructions, counted as 1 flop in "
                               "`achieved`): %g per datum x datum_evals_per_s" % hw_flop_per_datum)
        cores = os.cpu_count() or 1
        if world == 1 and not args.no_cpu_baseline:
            cpu_v, cpu_sample = cpu_baseline(cores)
            cpu = {"value": cpu_v, "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": cpu_sample}
        else:
            cpu = None
        others = None
        if world == 1 and not args.no_other_workloads:
            try:
                others = other_workloads(mb, torch, dev)
            except Exception as e:  # never let the side measurements take the headline line down
                others = {"error": repr(e)}
        line = {
            "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), lanes_per_chain=info["lanes_per_chain"],
                           chains_per_thread=info["chains_per_thread"], blocks=info["blocks"],
                           threads_per_block=info["threads_per_block"], smem_bytes=info["smem_bytes"]),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "note": "every e2e step uploads data + start points, runs the first 100 iterations of fresh chains "
                            "(stage-2 rate ~0.88 against ~0.71 in the steady state that `value` times) and downloads "
                            "theta/mean/cov/counters of every chain into pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "chains_with_error_status": status_bad,
            "other_workloads": others,
        }
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=1 << 20, help="chains per GPU")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the short C2/C4/C5 side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

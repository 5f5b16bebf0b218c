#!/usr/bin/env python
"""bench.py -- chain-steps/s of the batched adaptive-MH hot path on the BASELINE configurations.

  python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c4|c5|c1] [--scaling weak|strong]
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference (oracle/)

Default workload = BASELINE.json configs[2] ("C3", the config the whole-box target is quoted on):
2^20 independent DRAM+AM chains PER GPU (weak scaling) on the exponential-regression model with ndata = 10^4
observations staged in shared memory, drscale=2, adaptint=100, initcmatn=1, sigma2 Gibbs update on.  One bench "step"
= one pass of the hot path over every chain = `mcmcb_run(h, iters)` with iters = one adaptation interval (C3: ONE
kernel launch advancing every chain by 100 MCMC iterations).  The other workloads are the remaining BASELINE
configurations at their stated sizes:
  c2  4096 DRAM chains, 100-dim correlated Gaussian            (k2_step_kernel + k2_adapt_kernel)
  c4  65536 RAM chains, 50-dim banana, POOLED shape matrix: one allreduce pair per 100 steps inside the timed region
  c5  262144 SCAM chains, 200-parameter hierarchical model, POOLED rotation (one allreduce pair per 100 sweeps)
  c1  the reference's own testcase (11 data) batched as 2^22 chains
--scaling weak: the stated chain count per GPU; strong: the stated count in total, sharded over the GPUs.
Data are synthetic (seeded); FP64 throughout.

`value`  : device-timed (CUDA events on the kernels' stream, max over ranks), state resident in HBM.
`e2e`    : the same metric through the C ABI with HOST buffers: every step uploads every chain's start point
           (pinned) and the proposal covariance, runs the same number of iterations and downloads theta / mean /
           covariance / counters of every chain into pinned host buffers.  The start state is the steady state of the
           device-timed chains (their last points and their mean adapted covariance), so the work mix (stage-2 rate)
           is the one `value` times.
`roofline.traffic` and the hardware FP64 count come from the committed ncu capture profiles/r02_ncu_<workload>.json
(scripts/ncu_profile.py), which records the hash of the CUDA sources it was taken from: a capture of other sources is
reported as stale and its numbers are NOT used.
"""
import argparse
import glob
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SEED = 2024


def source_hash():
    """Identity of the CUDA sources the library is built from (profiles are keyed by it)."""
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "mcmcf90_b200", "csrc", "*.cu*")) +
                   glob.glob(os.path.join(ROOT, "mcmcf90_b200", "csrc", "*.h")) +
                   glob.glob(os.path.join(ROOT, "include", "*")))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def load_profile(workload):
    """Committed ncu capture of this workload's dominant kernel, or (None, reason)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_%s.json" % workload)
    if not os.path.exists(p):
        return None, "no capture committed (%s)" % os.path.relpath(p, ROOT)
    d = json.load(open(p))
    if d.get("source_hash") != source_hash():
        return None, "STALE capture: taken from sources %s, library sources are %s -- re-run scripts/ncu_profile.py" % (
            d.get("source_hash"), source_hash())
    return d, os.path.relpath(p, ROOT)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 7700.0 * 0.85}, "fallback of B200_PROFILING.md (no MEASURED_PEAKS.json)"


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE configuration as concrete synthetic inputs."""

    def __init__(self, name):
        self.name = name
        getattr(self, "_" + name)()

    def _c3(self):
        n = 10000
        rng = np.random.default_rng(SEED)
        x = 10.0 * np.arange(n) / (n - 1)
        y = 10.0 * np.exp(-0.1 * x) + rng.normal(0.0, np.sqrt(0.5), n)
        self.desc = "C3: DRAM+AM chains on exp-regression, ndata=%d in shared memory" % n
        self.model, self.oracle_model, self.blob_args = "expreg", "MODEL_EXPREG", ("blob_expreg", (x, y))
        self.d, self.ndata, self.chains, self.iters = 2, n, 1 << 20, 100
        self.nml = dict(adaptint=100, drscale=2.0, initcmatn=1, doburnin=0, burnintime=0, updatesigma=1, N0=1.0, S02=0.5)
        self.cmat0, self.sigma2, self.nobs = np.diag([0.2, 0.001]) * (11.0 / n), [0.5], [n]
        self.par0 = lambda nn, off: np.array([10.0, 0.1]) * (1.0 + 0.01 * np.random.default_rng([SEED, off]).standard_normal((nn, 2)))
        self.pool, self.bound = 0, "fp64"
        self.kernel = "k1_step_kernel"

    def _c1(self):
        x = np.arange(11.0)
        y = np.array([9.33, 9.40, 8.99, 7.06, 7.13, 6.69, 4.69, 4.24, 4.77, 3.86, 4.02])
        self.desc = "C1 batched: testcases/data.dat model (11 data), DRAM+AM"
        self.model, self.oracle_model, self.blob_args = "expreg", "MODEL_EXPREG", ("blob_expreg", (x, y))
        self.d, self.ndata, self.chains, self.iters = 2, 11, 1 << 22, 100
        self.nml = dict(adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0)
        self.cmat0, self.sigma2, self.nobs = np.diag([0.2, 0.001]), [0.5], [11]
        self.par0 = lambda nn, off: np.tile([10.0, 0.1], (nn, 1))
        self.pool, self.bound = 0, "fp64"
        self.kernel = "k1_step_kernel"

    def _c2(self):
        d, rho = 100, 0.9
        sd = 1.0 + 9.0 * np.arange(d) / (d - 1)
        sig = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d))) * np.outer(sd, sd)
        lam = np.linalg.inv(sig)
        self.desc = "C2: DRAM chains on a 100-dim correlated Gaussian target (private Cholesky factor per chain)"
        self.model, self.oracle_model, self.blob_args = "gauss", "MODEL_GAUSS", ("blob_gauss", (np.zeros(d), 0.5 * (lam + lam.T)))
        self.d, self.ndata, self.chains, self.iters = d, 0, 4096, 200
        self.nml = dict(adaptint=200, drscale=2.0, initcmatn=1, updatesigma=0)
        self.cmat0, self.sigma2, self.nobs = 0.01 * np.eye(d), [1.0], [1]
        self.par0 = lambda nn, off: np.zeros((nn, d))
        self.pool, self.bound = 0, "hbm"
        self.kernel = "k2_step_kernel"

    def _c4(self):
        d = 50
        self.desc = "C4: RAM chains on the 50-dim banana target, shape matrices pooled (averaged over ALL chains) every 100 steps"
        self.model, self.oracle_model, self.blob_args = "banana", "MODEL_BANANA", ("blob_banana", (d, 0.03))
        self.d, self.ndata, self.chains, self.iters = d, 0, 65536, 100
        self.nml = dict(method="ram", adaptint=100, updatesigma=0, alphatarget=0.234, nuparam=0.7)
        self.cmat0, self.sigma2, self.nobs = np.eye(d), [1.0], [1]
        self.par0 = lambda nn, off: np.zeros((nn, d))
        self.pool, self.bound = 1, "hbm"
        self.kernel = "k4_ram_step_kernel"

    def _c5(self):
        groups, per = 198, 10
        rng = np.random.default_rng(5)
        y = rng.normal(size=(groups, 1)) + rng.normal(size=(groups, per))
        d = groups + 2
        self.desc = ("C5: SCAM chains on the 200-parameter hierarchical model (1980 observations), ONE pooled rotation for all "
                     "chains rebuilt every 100 sweeps; one step = a sweep over the 200 components")
        self.model, self.oracle_model, self.blob_args = "hier", "MODEL_HIER", ("blob_hier", (y,))
        self.d, self.ndata, self.chains, self.iters = d, groups * per, 262144, 100
        self.nml = dict(method="scam", adaptint=100, initcmatn=1, updatesigma=0)
        self.cmat0, self.sigma2, self.nobs = 0.1 * np.eye(d), [1.0], [1]
        self.par0 = lambda nn, off: np.zeros((nn, d))
        self.pool, self.bound = 1, "fp64"
        self.kernel = "k5s_scam_step_kernel"

    def blob(self, mod):
        return getattr(mod, self.blob_args[0])(*self.blob_args[1])

    # ---- algorithmic work per chain-step, SURVEY.md 8d
    def flops_per_step(self, q, rows_per_step):
        d, n = self.d, self.ndata
        if self.name in ("c3", "c1"):
            f_prop, f_ss, f_q1, f_alpha = d * (d + 1), 6 * n, 4 * d * d + 6 * d, 12
            f_adapt = 5 * d * d * rows_per_step * self.nml["adaptint"] + d ** 3 / 3 + 2 * d ** 3 / 3
            return (1 + q) * (f_prop + f_ss + 3 * d) + q * f_q1 + (1 + q) * f_alpha + f_adapt / self.nml["adaptint"]
        if self.name == "c5":  # SCAM, minimal form: d component moves, each one full ssfunction
            return d * (2 * d + (6 * n + 5 * d) + 12)
        return None

    def hbm_bytes_per_step(self, q):
        d = self.d
        tri = d * (d + 1) // 2
        if self.name == "c2":
            return 8 * (tri * (1 + q) + 3 * d + 10)       # the factor is read once per proposal (1+q per step)
        if self.name == "c4":
            return 16 * tri + 8 * (3 * d + 10)             # RAM: the factor is read and rewritten every step
        return None

    def units_note(self):
        return {"c3": "F_step = (1+q)(d(d+1)+6n+3d) + q(4d^2+6d) + 12(1+q) + F_adapt/adaptint, exp counted as 1 flop",
                "c1": "as c3 with n = 11", "c5": "F_step = d (2d + 6n + 5d + 12), one sweep over d components",
                "c2": "bytes/step = 8 (d(d+1)/2 (1+q) + 3d + 10): private factor read once per proposal",
                "c4": "bytes/step = 16 d(d+1)/2 + 8 (3d + 10): private factor read and rewritten every step"}[self.name]


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


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def workload_config(W, args, chains_per_gpu, scaling):
    return {"workload": W.desc, "chains_per_gpu": int(chains_per_gpu), "mcmc_iterations_per_step": int(W.iters), "npar": W.d,
            "namelist": W.nml, "rng": "philox4x32-10", "scaling_mode": scaling,
            "parallelism": ("chains sharded over GPUs; pooled adaptation: 2 NCCL allreduces per %d iterations" % W.nml["adaptint"])
            if W.pool else "chains sharded over GPUs, no collective",
            "l2": "per-chain state exceeds L2 (no flush needed)" if W.name != "c2" else
                  "4096 private 40 KB factors = 164 MB exceed L2; 3-step warm-up"}


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_rate(W, cores, nch, nsimu, native):
    from oracle import oracle as O
    O.use_native(native)
    cfg = O.make_cfg(nsimu=nsimu, **W.nml)
    out = O.run_batch(cfg, getattr(O, W.oracle_model), W.blob(O), W.par0(nch, 0), W.cmat0, W.sigma2, W.nobs, seed=SEED,
                      chain0=0, nthreads=cores)
    return nch * (nsimu - 1) / max(out["seconds"], 1e-9), out["seconds"]


def cpu_baseline(W, cores, target_seconds=12.0, chains_per_core=4):
    """The oracle (C restatement of the reference, 'port') timed on the host cores: one chain per thread at a time
    (the reference is one chain per process).  Two builds: the reference's own flags (-O2, no -march,
    linux64.mk:38) and -O3 -march=native; the faster one is the reported value."""
    nch = cores * chains_per_core
    res = {}
    for native in (False, True):
        rate, _ = oracle_rate(W, cores, nch, 21, native)  # calibration
        nsimu = int(max(41, min(200000, 0.5 * target_seconds * rate / nch))) + 1
        rate, sec = oracle_rate(W, cores, nch, nsimu, native)
        res[native] = (rate, "%d chains x %d MCMC steps of %s, %d threads, %.1f s" % (nch, nsimu - 1, W.name.upper(), cores, sec))
    best = max(res, key=lambda k: res[k][0])
    return {"value": res[best][0], "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": res[best][1],
            "flags": "-O3 -march=native" if best else "-O2 -ffp-contract=off (reference flags, linux64.mk:38)",
            "value_reference_flags": res[False][0], "value_march_native": res[True][0],
            "note": ("C restatement of the reference (oracle/); no Fortran compiler on this box; pooled workloads: the CPU "
                     "chains adapt on their own (the reference has no pooling)") if W.pool else
                    "C restatement of the reference (oracle/); no Fortran compiler on this box"}


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    W = Workload(args.workload)
    cores = os.cpu_count() or 1
    from oracle import oracle as O
    O.use_native(True)
    nch = 8 * cores
    rate, _ = oracle_rate(W, cores, nch, 11, True)
    msteps = int(max(10, min(5000, 8.0 * rate / nch)))  # ~8 s per bench step
    cfg = O.make_cfg(nsimu=msteps + 1, **W.nml)
    par0, blob = W.par0(nch, 0), W.blob(O)
    times = []
    for it in range(args.warmup + args.steps):
        out = O.run_batch(cfg, getattr(O, W.oracle_model), blob, par0, W.cmat0, W.sigma2, W.nobs, seed=SEED + it, chain0=0,
                          nthreads=cores)
        if it >= args.warmup:
            times.append(out["seconds"])
    total = sum(times)
    value = nch * msteps * args.steps / total
    sample = "%d chains x %d MCMC steps per step on %d host threads (bounded sample of the %d-chain step)" % (
        nch, msteps, cores, W.chains)
    line = {
        "impl": "reference", "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(W, args, W.chains, args.scaling),
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "flags": "-O3 -march=native"},
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference (oracle/): the Fortran reference cannot be built here or on the GPU box "
                "(no Fortran compiler, profiles/r02_probe_fortran.txt)",
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ this framework
def run_ours(args):
    import torch
    import mcmcf90_b200 as mb
    from mcmcf90_b200 import parallel

    rank, world, local = dist_env()
    W = Workload(args.workload)
    if args.iters:
        W.iters = args.iters
    single_process = world == 1 and args.gpus > 1  # one handle drives all GPUs of the box (mcmcb_config.ngpus)
    G = args.gpus if single_process else 1
    ngpus_total = world * G
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)
    if args.chains:
        per_gpu = args.chains
    elif args.scaling == "strong":
        per_gpu = W.chains // ngpus_total
    else:
        per_gpu = W.chains
    N = per_gpu * G                   # chains of this process
    offset = rank * N
    blob = W.blob(mb.models)
    nsimu = 1 + W.iters * (2 * (args.warmup + args.steps) + 8)
    extra = dict(dump_stride=args.dump_stride) if args.dump_stride else {}
    cfg = mb.default_config(nchains=N, chain_offset=offset, seed=SEED, device=dev, nsimu=nsimu, model=W.model,
                            lanes_per_chain=args.lanes, pool_adapt=W.pool, ngpus=G, **W.nml, **extra)
    s = mb.Sampler(cfg)
    s.set_data(blob)
    par0_pinned = torch.empty((N, W.d), dtype=torch.float64).pin_memory()
    par0 = par0_pinned.numpy()
    par0[:] = W.par0(N, offset)
    s.set_initial(par0, W.cmat0, W.sigma2, W.nobs)
    nccl_world = False
    if world > 1 and W.pool:
        nccl_world = parallel.attach(s)
    streams = [torch.cuda.ExternalStream(s.stream_of(k), device=dev + k) for k in range(G)]

    def sync_all():
        for k in range(G):
            torch.cuda.synchronize(dev + k)

    def barrier():
        sync_all()
        if dist is not None:
            dist.barrier()
        sync_all()

    # FP64 pipe peak, measured on this GPU: burst (best 22 ms run) and sustained (2 s of back-to-back runs)
    burst = mb.dfma_peak(dev)[0]
    t_end, sus = time.time() + 2.0, []
    while time.time() < t_end:
        sus.append(mb.dfma_peak(dev)[0])
    sustained = float(np.mean(sus[len(sus) // 2:]))

    dumped = enq = 0

    def drain(keep=0):
        """The consumer of the streamed dumps: pops what is complete (host-side transposition included) and, when more than
        `keep` snapshots are still outstanding, waits for them -- a consumer that falls behind loses the oldest snapshots."""
        nonlocal dumped
        while args.dump_stride:
            if s.dump_pop_ex() is not None:
                dumped += 1
            elif enq - dumped > keep:
                time.sleep(0.0005)
            else:
                break

    def run_step():
        """One bench step = W.iters iterations of every chain.  With streamed dumps the host pops snapshots while the
        device runs on: the call is issued in pieces of at most 8 snapshots so that the ring never overflows."""
        if not args.dump_stride:
            s.run(W.iters, sync=False)
            return
        nonlocal enq
        left = W.iters
        while left > 0:
            n = min(left, args.dump_stride * 4)
            s.run(n, sync=False)
            enq += n // args.dump_stride
            drain(keep=4)   # at most 8 snapshots outstanding in a ring of up to 16
            left -= n

    for _ in range(args.warmup):
        run_step()
        drain()
    s.sync()
    drain()
    c0 = s.counters()
    l0, n0 = s.launches, s.nccl_calls
    clocks = ClockSampler(dev)
    barrier()
    clocks.start()

    def ev():
        return [torch.cuda.Event(enable_timing=True) for _ in range(G)]

    def rec(es):  # one event per device of the handle, each on that device's kernel stream
        for k in range(G):
            with torch.cuda.device(dev + k):
                es[k].record(streams[k])

    evs = [(ev(), ev()) for _ in range(args.steps)]
    e_all0, e_all1 = ev(), ev()
    rec(e_all0)
    for a, b in evs:
        rec(a)
        run_step()
        rec(b)
        drain()
    rec(e_all1)
    s.sync()
    drain()
    barrier()
    clk = clocks.stop()
    total_ms = max(e_all0[k].elapsed_time(e_all1[k]) for k in range(G))
    launch_ms = [max(a[k].elapsed_time(b[k]) for k in range(G)) for a, b in evs]
    launches = s.launches - l0
    nccl_calls = s.nccl_calls - n0
    c1 = s.counters()
    done_steps = W.iters * args.steps
    q = float((c1["drtries"] - c0["drtries"]).sum()) / (N * done_steps)
    rows = float((c1["chainind"] - c0["chainind"]).sum()) / (N * done_steps)
    stay = float((c1["stayed"] - c0["stayed"]).sum()) / (N * done_steps)
    status_bad = int((c1["status"] != 0).sum())
    info = s.info()

    # ---- end to end through the C ABI with host buffers, from the steady state of the chains above
    e2e = None
    if not args.no_e2e:
        d = W.d
        theta_end = s.fetch("par")
        if W.nml.get("method") == "ram":
            cov = W.cmat0  # RAM has no running covariance: the continuation starts from the initial shape
        elif W.pool:
            cov = s.pool_fetch()[2]  # the pooled covariance of the last tick
        else:
            cm = s.fetch("cmat")
            good = c1["status"] == 0
            cov = cm[good].mean(axis=0) if good.any() else W.cmat0
        cov = np.triu(cov) + np.triu(cov, 1).T
        s.close()
        s = None
        e2e_sampler = mb.Sampler(mb.default_config(nchains=N, chain_offset=offset, seed=SEED + 1, device=dev, nsimu=W.iters + 1,
                                                   model=W.model, lanes_per_chain=args.lanes, pool_adapt=W.pool, ngpus=G, **W.nml))
        if world > 1 and W.pool:
            parallel.attach(e2e_sampler)
        par0[:] = theta_end
        outs = {"par": torch.empty((N, d), dtype=torch.float64).pin_memory(),
                "mean": torch.empty((N, d), dtype=torch.float64).pin_memory(),
                "counters": torch.empty((N, 8), dtype=torch.int64).pin_memory()}
        if d <= 8:  # per-chain covariance of the small models; the large-npar workloads return the pooled one
            outs["cmat"] = torch.empty((N, d * d), dtype=torch.float64).pin_memory()
        h2d = par0.nbytes + blob.nbytes + cov.nbytes + 8 * len(W.sigma2) + 4 * len(W.nobs)
        e2e_steps = max(1, min(args.steps, 3))
        e2e_sampler.set_data(blob)
        e2e_sampler.set_initial(par0, cov, W.sigma2, W.nobs)  # allocation + first touch outside the timed region
        e2e_sampler.run(1)
        for w, t in outs.items():  # the library's device staging buffers for the downloads exist before the clock starts
            e2e_sampler.fetch(w, out=t.numpy())
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for it in range(e2e_steps):
            e2e_sampler.set_data(blob)                                   # H2D model data
            e2e_sampler.set_initial(par0, cov, W.sigma2, W.nobs)         # H2D start points (pinned) + init kernel
            e2e_sampler.run(W.iters, sync=False)
            out = [e2e_sampler.fetch(w, out=t.numpy()) for w, t in outs.items()]  # D2H results
            d2h = sum(o.nbytes for o in out)
            if W.pool:
                d2h += 8 * (1 + d + d * d)
                e2e_sampler.pool_fetch()
        sync_all()
        e2e_s = time.perf_counter() - t0
        ce = {k: outs["counters"].numpy()[:, i] for i, k in enumerate(mb.binding.COUNTER_NAMES)}
        q_e2e = float(ce["drtries"].sum()) / (N * W.iters)
        e2e_sampler.close()
        e2e = (e2e_s, e2e_steps, h2d, d2h, q_e2e)

    t_total = torch.tensor([total_ms, (e2e[0] * 1e3) if e2e else 0.0], dtype=torch.float64, device="cuda:%d" % dev)
    if dist is not None:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = [float(v) for v in t_total.cpu()]
    value = world * N * done_steps / (total_ms_max * 1e-3)

    if rank == 0:
        avg_launch_ms = float(np.mean(launch_ms))
        rate_gpu = per_gpu * W.iters / (avg_launch_ms * 1e-3)        # chain-steps/s of one GPU inside one step
        prof, prof_src = load_profile(W.name)
        peaks, peaks_src = measured_peaks()
        roof = {"bound": W.bound, "units": W.units_note(), "stage2_rate_q": q, "accept_rate": 1 - stay,
                "kernel": W.kernel, "avg_launch_ms": avg_launch_ms, "launch_ms": launch_ms,
                "profile": prof_src}
        if W.bound == "fp64":
            fl = W.flops_per_step(q, rows)
            achieved = rate_gpu * fl / 1e12
            roof.update({"achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                         "peak_source": "in-bench DFMA microbenchmark on this GPU, sustained 2 s (burst %.2f); MEASURED_PEAKS.json "
                                        "holds only HBM and bf16-tensor peaks, neither bounds this kernel" % burst,
                         "algorithmic_flops_per_chain_step": fl})
            if W.ndata:
                roof["datum_evals_per_s"] = rate_gpu * (1 + q) * W.ndata * (W.d if W.name == "c5" else 1)
            if W.name == "c3":  # round 1's definition, for continuity: the 9 FP64 instructions (16 flops) of the datum loop only
                roof["frac_hw_datum_loop"] = roof["datum_evals_per_s"] * 16 / 1e12 / sustained
        else:
            by = W.hbm_bytes_per_step(q)
            achieved = rate_gpu * by / 1e9
            roof.update({"achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "peak_source": peaks_src, "algorithmic_bytes_per_chain_step": by})
        # measured hardware counts of ONE launch of the dominant kernel, from the committed ncu capture
        roof["traffic"] = None
        if prof:
            per_launch_units = prof["chains"] * prof["iterations"]
            scale = per_gpu * W.iters / per_launch_units   # the capture may cover fewer iterations than a bench step
            roof["traffic"] = prof["dram_bytes"] * scale
            roof["traffic_unit"] = "bytes per bench step (ncu dram__bytes_read+write of one launch, scaled by chain-steps)"
            hwf = prof.get("fp64_flops")
            if hwf and W.bound == "fp64":
                hw = hwf / per_launch_units * rate_gpu / 1e12
                roof["achieved_hw"] = hw
                roof["frac_hw"] = hw / sustained
                roof["hw_note"] = ("FP64 flops the kernel executes per chain-step (ncu smsp__sass_thread_inst_executed_op_"
                                   "dfma x2 + dadd + dmul of the captured launch: %.4g) x chain-steps/s" % (hwf / per_launch_units))
            for k in ("fp64_pipe_pct", "issue_active_pct", "shared_wavefronts", "shared_bank_conflicts", "duration_ms"):
                if k in prof:
                    roof["ncu_" + k] = prof[k]
        if W.name in ("c3", "c1"):
            roof["algorithmic_bytes_per_launch"] = 2.0 * 212 * per_gpu
        cores = os.cpu_count() or 1
        cpu = None
        if ngpus_total == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(W, cores)
        line = {
            "metric": "chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "n_gpus": ngpus_total,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(W, args, per_gpu, args.scaling), processes=world, gpus_per_process=G,
                           source_hash=source_hash(), lanes_per_chain=info["lanes_per_chain"],
                           chains_per_thread=info["chains_per_thread"], blocks=info["blocks"],
                           threads_per_block=info["threads_per_block"], smem_bytes=info["smem_bytes"]),
            "clocks": clk,
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "chains_with_error_status": status_bad,
        }
        if W.pool:
            line["collective"] = {"pooled_ticks_in_timed_region": int(args.steps * W.iters // W.nml["adaptint"]),
                                  "allreduce_doubles_per_tick": 1 + W.d + W.d * W.d,
                                  "backend": "NCCL in-process (ncclCommInitAll), %d calls" % nccl_calls if G > 1 else
                                             ("NCCL via torch.distributed callback" if nccl_world else "single GPU: local sums")}
        if e2e:
            e2e_s, e2e_steps, h2d, d2h, q_e2e = e2e
            line["e2e"] = {"value": world * N * W.iters * e2e_steps / (e2e_ms_max * 1e-3), "unit": "chain-steps/s",
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                           "stage2_rate_q": q_e2e,
                           "note": "every e2e step uploads the model data, every chain's start point (the steady-state points of the "
                                   "device-timed chains) and the adapted proposal covariance, runs %d iterations and downloads "
                                   "theta/mean/cov/counters of every chain into pinned host buffers" % W.iters}
        if args.dump_stride:
            line["dumps"] = {"dump_stride": args.dump_stride, "snapshots_popped": dumped, "snapshots_enqueued": enq,
                             "bytes_per_snapshot": int(N * (W.d + 2 * len(W.sigma2) + 1) * 8),
                             "note": "theta/ss/sigma2 of every chain streamed to pinned host buffers on a copy stream and popped by "
                                     "the host inside the timed region; compare `value` with a run without --dump-stride"}
        print(json.dumps(line))
    if s is not None:
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
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5", "c1"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--chains", type=int, default=0, help="override: chains per GPU")
    ap.add_argument("--iters", type=int, default=0, help="override: MCMC iterations per bench step")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--dump-stride", type=int, default=0, help="stream theta/ss/sigma2 of every chain to the host every n iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="accepted for compatibility (no effect)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

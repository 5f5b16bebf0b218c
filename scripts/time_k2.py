"""Single-GPU timing of the large-npar kernel on BASELINE configs C2 / C4 shapes (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcf90_b200 as mb


def gauss_target(d, rho=0.9):
    s = 1.0 + 9.0 * np.arange(d) / max(d - 1, 1)
    Sig = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d))) * np.outer(s, s)
    lam = np.linalg.inv(Sig)
    return np.zeros(d), 0.5 * (lam + lam.T)


def timeit(name, cfg_kw, model, blob, d, N, steps, cmat0, bytes_per_step):
    s = mb.Sampler(mb.default_config(nchains=N, seed=12345, model=model, **cfg_kw))
    s.set_data(blob)
    s.set_initial(np.zeros(d), cmat0, [1.0], [1])
    s.run(steps)
    st = torch.cuda.ExternalStream(s.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        c0 = s.counters()
        l0 = s.launches
        e0.record(st)
        s.run(steps, sync=False)
        e1.record(st)
        s.sync()
        ms = e0.elapsed_time(e1)
        c1 = s.counters()
        q = (c1["drtries"] - c0["drtries"]).sum() / (N * steps)
        acc = 1 - (c1["stayed"] - c0["stayed"]).sum() / (N * steps)
        rate = N * steps / ms * 1e3
        b = bytes_per_step(q)
        print("%s d=%d N=%d steps=%d: %.2f ms -> %.3e chain-steps/s q=%.3f acc=%.3f  alg %.0f B/step -> %.1f GB/s  launches=%d bad=%d info=%s" % (
            name, d, N, steps, ms, rate, q, acc, b, rate * b / 1e9, s.launches - l0, (c1["status"] != 0).sum(), s.info()))
    s.close()


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("c2", "all"):
    d, N = 100, int(os.environ.get("N_C2", 4096))
    mu, lam = gauss_target(d)
    T = d * (d + 1) // 2
    timeit("C2 gauss DRAM", dict(nsimu=100000, adaptint=200, drscale=2.0, initcmatn=1, updatesigma=0), "gauss",
           mb.models.blob_gauss(mu, lam), d, N, 200, 0.01 * np.eye(d), lambda q: 8 * (T * (1 + q) + 3 * d + 10))
if which in ("c4", "all"):
    d, N = 50, int(os.environ.get("N_C4", 65536))
    T = d * (d + 1) // 2
    timeit("C4 banana RAM", dict(method=mb.RAM if hasattr(mb, "RAM") else 1, nsimu=100000, updatesigma=0,
                                 alphatarget=0.234, nuparam=0.7), "banana",
           mb.models.blob_banana(d, 0.03), d, N, 100, np.eye(d), lambda q: 16 * T + 8 * (3 * d + 10))
if which in ("c5", "all"):
    G, J, N = 198, 10, int(os.environ.get("N_C5", 2048))
    d = G + 2
    rng = np.random.default_rng(5)
    y = rng.normal(size=(G, 1)) + rng.normal(size=(G, J))
    # algorithmic HBM bytes per sweep: every component move reads one column of U (d doubles)
    timeit("C5 hier SCAM", dict(method=mb.SCAM, nsimu=100000, adaptint=100, initcmatn=1, updatesigma=0), "hier",
           mb.models.blob_hier(y), d, N, 20, 0.1 * np.eye(d), lambda q: 8 * (d * d + 3 * d + 10))

"""Where the end-to-end step of bench.py spends its time (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcf90_b200 as mb
import bench as B
N = 1 << 20
x, y = B.synth_data()
blob = mb.models.blob_expreg(x, y)
par0_t = torch.empty((N, 2), dtype=torch.float64).pin_memory(); par0 = par0_t.numpy(); par0[:] = B.start_points(N, 0)
s = mb.Sampler(mb.default_config(nchains=N, seed=1, nsimu=101, model="expreg", **B.NML))
outs = {"par": torch.empty((N, 2), dtype=torch.float64).pin_memory(), "mean": torch.empty((N, 2), dtype=torch.float64).pin_memory(),
        "cmat": torch.empty((N, 4), dtype=torch.float64).pin_memory(), "counters": torch.empty((N, 8), dtype=torch.int64).pin_memory()}
def T(label, f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); print("%-14s %8.2f ms" % (label, 1e3 * (time.perf_counter() - t))); return r
for it in range(3):
    print("-- step", it)
    T("set_data", lambda: s.set_data(blob))
    T("set_initial", lambda: s.set_initial(par0, B.CMAT0, [0.5], [B.NDATA]))
    T("run", lambda: s.run(100))
    for w, t in outs.items():
        T("fetch " + w, lambda: s.fetch(w, out=t.numpy()))
    T("fetch pageable", lambda: s.fetch("counters"))

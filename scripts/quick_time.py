"""Quick single-GPU timing of the C3 workload (dev tool, not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcf90_b200 as mb
from tests import cases

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ndata = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
print("dfma peak TFLOP/s, ms:", mb.dfma_peak(0))
x, y = cases.synth_expreg(ndata)
blob = mb.models.blob_expreg(x, y)
nml = dict(nsimu=100000, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.5)
par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(0).normal(size=(N, 2)))
s = mb.Sampler(mb.default_config(nchains=N, seed=1, lanes_per_chain=lanes, **nml))
s.set_data(blob)
s.set_initial(par0, cases.CMAT0 * (11.0 / ndata), [0.5], [ndata])
s.run(steps)  # warm-up incl. initial evaluation
st = torch.cuda.ExternalStream(s.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    c0 = s.counters()
    e0.record(st)
    s.run(steps, sync=False)
    e1.record(st)
    s.sync()
    ms = e0.elapsed_time(e1)
    c1 = s.counters()
    q = (c1["drtries"] - c0["drtries"]).sum() / (N * steps)
    print("N=%d steps=%d lanes=%s: %.2f ms -> %.3e chain-steps/s, q=%.3f, datum-evals/s=%.3e  info=%s" % (
        N, steps, s.info()["lanes_per_chain"], ms, N * steps / ms * 1e3, q, N * steps * (1 + q) * ndata / ms * 1e3, s.info()))
print("accept rate", 1 - c1["stayed"].mean() / c1["simuind"].mean())

"""Early-rejection sampler on the C3 shape: throughput with and without the warp-vote early exit of the data loop
(MCMCB_ER_NOEXIT=1), against plain AM (drscale = 0) -- dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcf90_b200 as mb
from tests import cases

N, steps, ndata = 1 << 20, 50, 10000
x, y = cases.synth_expreg(ndata)
blob = mb.models.blob_expreg(x, y)
par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(0).normal(size=(N, 2)))
for label, kw, env in (("AM (drscale=0)", dict(method="dram", drscale=0.0), "0"), ("ER, batched", dict(method="er"), "0"),
                       ("ER, early exit", dict(method="er"), "1")):
    os.environ["MCMCB_ER_EXIT"] = env
    s = mb.Sampler(mb.default_config(nchains=N, seed=1, lanes_per_chain=1, nsimu=100000, adaptint=100, initcmatn=1,
                                     updatesigma=1, N0=1.0, S02=0.5, **kw))
    s.set_data(blob)
    s.set_initial(par0, cases.CMAT0 * (11.0 / ndata), [0.5], [ndata])
    s.run(200)  # past the first adaptation
    st = torch.cuda.ExternalStream(s.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = s.counters()
    e0.record(st); s.run(steps, sync=False); e1.record(st); s.sync()
    c1 = s.counters()
    ms = e0.elapsed_time(e1)
    acc = 1 - (c1["stayed"] - c0["stayed"]).sum() / (N * steps)
    print("%-16s %8.1f ms  %.3e chain-steps/s  accept %.3f  chains/thread %d" % (label, ms, N * steps / ms * 1e3, acc, s.info()["chains_per_thread"]))
    s.close()

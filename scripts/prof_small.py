"""Small C3-shaped run for ncu captures (one resident wave, few MCMC steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcmcf90_b200 as mb
from tests import cases
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x, y = cases.synth_expreg(10000)
blob = mb.models.blob_expreg(x, y)
nml = dict(nsimu=100000, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.5)
s = mb.Sampler(mb.default_config(nchains=N, seed=1, **nml))
s.set_data(blob)
s.set_initial(cases.PAR0, cases.CMAT0 * (11.0 / 10000), [0.5], [10000])
s.run(steps)
s.run(steps)
print(s.counters()["simuind"][:3])

import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import mcmcf90_b200 as mb
from tests import cases
N = 1 << 21
blob = mb.models.blob_expreg(cases.DATA_X, cases.DATA_Y)
nml = dict(nsimu=100000, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0)
s = mb.Sampler(mb.default_config(nchains=N, seed=1, **nml))
s.set_data(blob)
s.set_initial(cases.PAR0, cases.CMAT0, [0.5], [11])
s.run(100)
s.run(50)
print(s.counters()["simuind"][:3])

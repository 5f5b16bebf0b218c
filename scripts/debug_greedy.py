"""Debug aid: device vs oracle at successive burn-in ticks (greedy / AP configurations)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import mcmcf90_b200 as mb
from oracle import oracle as O
from tests import cases
from tests.test_k1_parity import BLOB11

nml = dict(nsimu=401, adaptint=50, burnintime=300, doburnin=1, badaptint=25, greedy=1, scalelimit=0.05,
           drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0)
nml.update(eval(sys.argv[1]) if len(sys.argv) > 1 else {})
N = 2
u = np.random.default_rng(31).random((N, 40 * 901))
cfg = mb.default_config(nchains=N, store_chains=-1, model="expreg", rng_mode=mb.RNG_INJECTED, **nml)
s = mb.Sampler(cfg); s.set_data(BLOB11); s.set_initial(cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS); s.inject_uniforms(u)
ch = O.Chain(O.make_cfg(**nml), O.MODEL_EXPREG, BLOB11, cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS); ch.inject(u[0])
i = 1
while i < nml["nsimu"]:
    nxt = min(i + 5, nml["nsimu"])
    s.run(nxt - i); ch.advance(nxt); i = nxt
    c = s.counters(); r = ch.results()
    R = s.fetch("R")[0]
    print(i, "dev stayed", c["stayed"][0], "orc", r["stayed"], "nd", c["ndrawn"][0], r["ndrawn"], "st", c["status"][0], r["status"],
          "R dev", R[np.triu_indices(2)], "orc", r["R"][np.triu_indices(2)], "wsum", s.fetch("wsum")[0, 0], r["wsum"])
    if c["ndrawn"][0] != r["ndrawn"]:
        break

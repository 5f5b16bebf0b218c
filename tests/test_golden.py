"""Committed golden vectors (tests/golden/golden_c1.npz, made by tests/golden/make_golden.py from the
oracle on the reference's shipped testcase inputs under injected uniform streams).

not gpu: the oracle still reproduces them (counters exactly, values to 1e-13: a different libm may move the
         last bit of exp/log/pow);
gpu:     the CUDA path through the C ABI reproduces them (counters exactly; values at the tolerances of the
         parity tests: 1e-10 DRAM/AM, 1e-9 RAM whose dchdd amplifies rounding, 1e-8 SCAM)."""
import os

import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O
from tests.golden import make_golden as G

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_c1.npz"))
CNT = ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "status", "ndrawn")


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(name):
    model_id, blob, par0, cmat0, sigma2, nobs = G.inputs(name)
    ch = O.Chain(O.make_cfg(**G.CASES[name]), model_id, blob, par0, cmat0, sigma2, nobs)
    ch.inject(G.uniforms(name))
    ch.run()
    r = ch.results()
    assert [r[k] for k in CNT] == list(GOLD[name + "_counters"])
    assert np.array_equal(r["chain"][:, -1], GOLD[name + "_chain"][:, -1])  # run-length column
    for k in ("chain", "sschain", "s2chain", "R", "cmat", "mean", "par", "sigma2"):
        np.testing.assert_allclose(r[k], GOLD["%s_%s" % (name, k)], rtol=1e-13, atol=1e-300, err_msg=k)


def test_golden_run_lengths_are_consistent():
    # size-independent property of MCMC_savechain (MCMC_aux.F90:166-185): the repeat counts of the stored
    # rows add up to the number of simulated steps, and rows - 1 == accepted steps
    for name in G.CASES:
        c = dict(zip(CNT, GOLD[name + "_counters"]))
        assert GOLD[name + "_chain"][:, -1].sum() == c["simuind"] == G.NSIMU
        assert c["chainind"] - 1 == (G.NSIMU - 1) - c["stayed"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,rtol", [("shipped", 1e-10), ("dram", 1e-10), ("ram", 1e-9), ("scam", 1e-8), ("scam_hier", 1e-8),
                                       ("er", 1e-10), ("ap", 1e-9), ("greedy", 1e-9),
                                       ("gauss_dram", 1e-9), ("gauss_ram", 1e-8), ("gauss_er", 1e-9), ("gauss_ap", 1e-9),
                                       ("gauss_greedy", 1e-9)])
def test_cuda_path_reproduces_golden(name, rtol):
    # the 6-dimensional Gaussian cases run on BOTH kernel families: "gauss" has a compile-time-npar registration for the
    # register kernel at npar = 6 (the automatic choice) and the run-time-npar one for the warp-per-chain kernels
    for kernel in ((0, 2) if name.startswith("gauss_") else (0,)):
        _golden_case(name, rtol, kernel)


def _golden_case(name, rtol, kernel):
    model_id, blob, par0, cmat0, sigma2, nobs = G.inputs(name)
    model = {O.MODEL_EXPREG: "expreg", O.MODEL_HIER: "hier", O.MODEL_GAUSS: "gauss"}[model_id]
    N = 3  # the same stream in every chain: all chains must reproduce the golden run
    cfg = mb.default_config(nchains=N, store_chains=-1, model=model, rng_mode=mb.RNG_INJECTED, kernel=kernel, **G.CASES[name])
    s = mb.Sampler(cfg)
    s.set_data(blob)
    s.set_initial(np.asarray(par0, dtype=float), cmat0, sigma2, nobs)
    s.inject_uniforms(np.tile(G.uniforms(name), (N, 1)))
    s.run(G.NSIMU - 1)
    if name.startswith("gauss_"):
        assert s.info()["kernel"] == (2 if kernel == 2 else 1)
    cnt = s.counters()
    par = s.fetch("par")
    for c in range(N):
        assert [int(cnt[k][c]) for k in CNT] == list(GOLD[name + "_counters"]), c
        g = s.fetch_chain(c)
        ref = GOLD[name + "_chain"]
        assert np.array_equal(g["chain"][:, -1], ref[:, -1])
        scale = np.abs(ref[:, :-1]).max()
        np.testing.assert_allclose(g["chain"][:, :-1], ref[:, :-1], rtol=rtol, atol=rtol * scale)
        np.testing.assert_allclose(g["sschain"][:, :-1], GOLD[name + "_sschain"][:, :-1], rtol=max(rtol, 1e-9))
        np.testing.assert_allclose(par[c], GOLD[name + "_par"], rtol=rtol, atol=rtol * scale)
    s.close()

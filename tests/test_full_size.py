"""BASELINE configurations C2, C4 and C5 at their FULL population sizes (4096 / 65 536 / 262 144 chains): parity of
sampled chains with the oracle where the chains are still the reference's own (before the first pooled tick), and
size-independent invariants everywhere (counter identities, status words, shared pooled factor).  The C3 population (2^20
chains) is covered by tests/test_k1_parity.py::test_full_size_properties_one_million_chains and
tests/test_r02_coverage.py::test_c3_four_chains_per_thread_four_ticks_with_dr."""
import numpy as np
import pytest

import bench
import mcmcf90_b200 as mb
from oracle import oracle as O

pytestmark = pytest.mark.gpu
CNT = ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "ndrawn")


def invariants(cnt, steps, dr):
    assert (cnt["status"] == 0).all()
    assert (cnt["simuind"] == steps + 1).all()
    assert (cnt["chainind"] - 1 == steps - cnt["stayed"]).all()      # MCMC_savechain: a row per accepted step
    if dr:
        assert (cnt["drtries"] >= cnt["stayed"]).all() and (cnt["draccepted"] <= cnt["drtries"]).all()
    else:
        assert (cnt["drtries"] == 0).all()
    assert (cnt["ndrawn"] > 0).all()


def sampler(W, N, nsimu, **extra):
    s = mb.Sampler(mb.default_config(nchains=N, seed=bench.SEED, nsimu=nsimu, model=W.model, pool_adapt=W.pool, **dict(W.nml, **extra)))
    s.set_data(W.blob(mb.models))
    return s


def oracle_chain(W, nsimu, par0, c, **extra):
    ch = O.Chain(O.make_cfg(nsimu=nsimu, **dict(W.nml, **extra)), getattr(O, W.oracle_model), W.blob(O), par0, W.cmat0, W.sigma2, W.nobs)
    ch.philox(bench.SEED, c)
    ch.run()
    return ch.results()


def test_c2_full_population_two_ticks():
    W = bench.Workload("c2")   # adaptint = 200: ~175 accepted rows per interval, enough for a 100 x 100 covariance of full rank
    N, steps = W.chains, 400
    rng = np.random.default_rng(2)
    par0 = 0.05 * rng.normal(size=(N, W.d))
    s = sampler(W, N, steps + 1)
    s.set_initial(par0, W.cmat0, W.sigma2, W.nobs)
    s.run(steps)
    cnt, par = s.counters(), s.fetch("par")
    assert (cnt["simuind"] == steps + 1).all() and (cnt["chainind"] - 1 == steps - cnt["stayed"]).all()
    assert (cnt["drtries"] >= cnt["stayed"]).all() and (cnt["draccepted"] <= cnt["drtries"]).all()
    for c in (0, 2047, N - 1):
        r = oracle_chain(W, steps + 1, par0[c], c)
        for k in CNT + ("status",):  # a chain whose covariance is not yet of full rank at a tick keeps its old factor
            assert cnt[k][c] == r[k], (c, k)   # (MCMC_adapt.F90:169-171): flagged by both, the same chains
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8, atol=1e-10)
    assert (cnt["status"] == 0).mean() > 0.9
    s.close()


def test_c4_full_population_pooled_ram():
    W = bench.Workload("c4")
    N, a = W.chains, 20
    s = sampler(W, N, 2 * a + 2, adaptint=a)
    rng = np.random.default_rng(4)
    par0 = 0.1 * rng.normal(size=(N, W.d))
    s.set_initial(par0, W.cmat0, W.sigma2, W.nobs)
    s.run(a - 2)   # stops one step short of the first pooled tick: every chain is still the reference's own RAM chain
    assert s.info()["lanes_per_chain"] == 1  # the thread-per-chain kernel
    cnt, par = s.counters(), s.fetch("par")
    invariants(cnt, a - 2, dr=False)
    for c in (0, 31, N - 1):
        r = oracle_chain(W, a - 1, par0[c], c, adaptint=a)
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k)
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(np.triu(s.fetch_stats(c)["R"]), np.triu(r["R"]), rtol=1e-7, atol=1e-10)
    s.run(a + 3)  # across two pooled ticks (simuind = 20, 40)
    cnt = s.counters()
    invariants(cnt, 2 * a + 1, dr=False)
    Wt, _, S = s.pool_fetch()
    assert Wt == N and np.all(np.linalg.eigvalsh(0.5 * (S + S.T)) > 0)
    s.close()


def test_c5_full_population_pooled_scam():
    W = bench.Workload("c5")
    N, a = W.chains, 3
    s = sampler(W, N, 2 * a + 2, adaptint=a)
    rng = np.random.default_rng(5)
    sub = 0.1 * rng.normal(size=(3, W.d))
    par0 = np.zeros((N, W.d))
    par0[[0, 1000, N - 1]] = sub
    s.set_initial(par0, W.cmat0, W.sigma2, W.nobs)
    s.run(a - 1)          # sweeps before the first pooled tick: the reference's own SCAM chains
    assert s.info()["lanes_per_chain"] == 4 and s.info()["smem_bytes"] > 200 * 1024   # theta in shared memory, k5s_scam.cuh
    cnt = s.counters()
    invariants(cnt, a - 1, dr=False)
    for c in (0, 1000, N - 1):
        r = oracle_chain(W, a, par0[c], c, adaptint=a)
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k)
        st = s.fetch_stats(c)
        assert st["counters"]["ndrawn"] == r["ndrawn"]
    par = s.fetch("par")
    for c in (0, 1000, N - 1):
        r = oracle_chain(W, a, par0[c], c, adaptint=a)
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8, atol=1e-10)
    del par
    s.run(a + 2)          # two pooled ticks (covariance of 262 144 chains merged, one SVD, shared rotation)
    cnt = s.counters()
    invariants(cnt, 2 * a + 1, dr=False)
    Wt, mu, cov = s.pool_fetch()
    assert Wt > N and np.isfinite(cov).all() and np.all(np.diag(cov) > 0)
    q = s.fetch_stats(N - 1)
    assert np.isfinite(q["R"]).all()
    s.close()

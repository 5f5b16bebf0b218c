"""GPU parity of the large-npar warp-per-chain kernel (K2) against the CPU oracle.

K2 differs from the reference's arithmetic at rounding level only (summation order of the
triangular product, (R'z)/drscale instead of a stored R2, matrix-free DR ratio), so accept
counts / chain indices must still be bit-exact under injected draws while values are
compared at 1e-10 (they accumulate ~d*eps per step instead of K1's ~eps)."""
import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def gauss_target(d, rho=0.9, seed=0):
    s = 1.0 + 9.0 * np.arange(d) / max(d - 1, 1)
    Sig = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d))) * np.outer(s, s)
    lam = np.linalg.inv(Sig)
    lam = 0.5 * (lam + lam.T)
    return np.zeros(d), lam, Sig


def run_gpu(nml, N, model, blob, par0, cmat0, u=None, seed=0, splits=None, chain_offset=0, sigma2=(1.0,), nobs=(1,),
            prior=None):
    # kernel=2: these are the parity tests of the warp-per-chain kernels ("gauss" at npar = 3..6, 8 would otherwise take
    # its compile-time-npar registration for the register kernel, tests/test_r02_coverage.py)
    cfg = mb.default_config(nchains=N, seed=seed, store_chains=-1, model=model, chain_offset=chain_offset, kernel=2,
                            rng_mode=mb.RNG_INJECTED if u is not None else mb.RNG_PHILOX, **nml)
    s = mb.Sampler(cfg)
    s.set_data(blob)
    if prior is not None:
        s.set_priors(*prior)
    s.set_initial(par0, cmat0, list(sigma2), list(nobs))
    if u is not None:
        s.inject_uniforms(u)
    for n in (splits or [nml["nsimu"] - 1]):
        s.run(n)
    return s


def run_oracle(nml, k, model_id, blob, par0, cmat0, u=None, seed=0, chain_offset=0, sigma2=(1.0,), nobs=(1,), prior=None):
    ch = O.Chain(O.make_cfg(**nml), model_id, blob, par0, cmat0, list(sigma2), list(nobs), prior=prior)
    if u is not None:
        ch.inject(u[k])
    else:
        ch.philox(seed, chain_offset + k)
    ch.run()
    r = ch.results()
    r["erstayed"] = ch.counters()["erstayed"]
    return r


def compare(s, nml, ks, model_id, blob, par0, cmat0, u=None, seed=0, chain_offset=0, rtol=RTOL, sigma2=(1.0,),
            nobs=(1,), at_tick=None):
    cnt = s.counters()
    par, ss, s2 = s.fetch("par"), s.fetch("ss"), s.fetch("sigma2")
    R, cm, mean, wsum = s.fetch("R"), s.fetch("cmat"), s.fetch("mean"), s.fetch("wsum")
    d = par.shape[1]
    iu = np.triu_indices(d)
    for k in ks:
        p0 = par0 if np.ndim(par0) == 1 else par0[k]
        r = run_oracle(nml, k, model_id, blob, p0, cmat0, u=u, seed=seed, chain_offset=chain_offset, sigma2=sigma2,
                       nobs=nobs)
        for key in ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "status", "ndrawn"):
            assert cnt[key][k] == r[key], (k, key)
        g = s.fetch_chain(k)
        assert np.array_equal(g["chain"][:, -1], r["chain"][:, -1])
        scale = np.abs(r["chain"][:, :-1]).max()  # components cross zero: absolute floor relative to the chain's scale
        np.testing.assert_allclose(g["chain"][:, :-1], r["chain"][:, :-1], rtol=rtol, atol=rtol * scale)
        np.testing.assert_allclose(g["sschain"][:, :-1], r["sschain"][:, :-1], rtol=max(rtol, 1e-9))
        np.testing.assert_allclose(g["s2chain"][:r["simuind"]], r["s2chain"][:r["simuind"]], rtol=max(rtol, 1e-9))
        np.testing.assert_allclose(par[k], r["par"], rtol=rtol, atol=rtol * scale)
        sc = np.abs(r["R"][iu]).max()
        np.testing.assert_allclose(R[k][iu], r["R"][iu], rtol=1e-7, atol=1e-9 * sc)
        ns, ai = nml["nsimu"], nml.get("adaptint", 100)
        tick = (nml.get("method", "dram") == "dram" and nml.get("doadapt", 1) and ns % ai == 0
                and ns >= nml.get("burnintime", 0) + ai) if at_tick is None else at_tick
        if tick:
            assert wsum[k, 0] == r["wsum"]
            np.testing.assert_allclose(mean[k], r["mean"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(cm[k][iu], r["cmat"][iu], rtol=1e-7, atol=1e-9 * np.abs(r["cmat"]).max())


@pytest.mark.parametrize("d", [3, 12, 40])
@pytest.mark.parametrize("variant", ["dram", "am", "burnin"])
def test_gauss_dram_injected_counts_bit_exact(d, variant):
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = {"dram": dict(nsimu=600, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=0),
           "am": dict(nsimu=500, adaptint=50, drscale=0.0, initcmatn=3, updatesigma=1, N0=4.0, S02=1.0),
           "burnin": dict(nsimu=600, adaptint=100, burnintime=200, doburnin=1, badaptint=40, drscale=3.0,
                          initcmatn=1, scalelimit=0.3, updatesigma=0)}[variant]
    N = 5
    u = np.random.default_rng(d).random((N, (4 * d + 40) * nml["nsimu"]))
    par0 = np.zeros(d)
    cmat0 = np.eye(d) * 0.5
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u)
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u)
    s.close()


@pytest.mark.parametrize("d", [3, 40])
def test_gauss_early_rejection_sampler(d):
    # method 'er' (MCMC_run_er.F90:12-107) on the warp-per-chain kernel, with a Gaussian prior so that the
    # prior-only rejections (erstayed) occur
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(method="er", nsimu=500, adaptint=100, initcmatn=3, updatesigma=1, N0=4.0, S02=1.0, drscale=2.0)
    N = 4
    u = np.random.default_rng(100 + d).random((N, (4 * d + 40) * nml["nsimu"]))
    par0, cmat0 = np.zeros(d), np.eye(d) * 0.5
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u, splits=[123, 376])
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u, at_tick=True)
    assert (s.counters()["drtries"] == 0).all()
    s.close()


@pytest.mark.parametrize("d,variant", [(12, "dram"), (40, "burnin"), (100, "dram"), (70, "er")])
def test_group_of_warps_per_chain_kernel(d, variant, monkeypatch):
    # the opt-in kernel of csrc/k2g_group.cuh (MCMCB_K2_GROUP=1): several warps per chain, factor packed in shared
    # memory; same parity bars as the warp-per-chain kernel
    monkeypatch.setenv("MCMCB_K2_GROUP", "1")
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = {"dram": dict(nsimu=400, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=0),
           "burnin": dict(nsimu=400, adaptint=100, burnintime=200, doburnin=1, badaptint=40, drscale=3.0,
                          initcmatn=1, scalelimit=0.3, updatesigma=1, N0=4.0, S02=1.0),
           "er": dict(method="er", nsimu=400, adaptint=100, initcmatn=3, updatesigma=0)}[variant]
    N = 7
    u = np.random.default_rng(1000 + d).random((N, (4 * d + 40) * nml["nsimu"]))
    par0, cmat0 = np.zeros(d), np.eye(d) * (0.5 if d < 50 else 0.01)
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u, splits=[150, 249])
    assert s.info()["lanes_per_chain"] == min(256, (d + 31) // 32 * 32)
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u, rtol=1e-9 if d >= 70 else RTOL)
    s.close()


@pytest.mark.parametrize("d,group", [(3, 0), (12, 0), (12, 1)])
@pytest.mark.parametrize("variant", ["ap", "ap_short", "ap_burnin", "greedy", "greedy_no_reset", "greedy_only"])
def test_ap_windows_and_greedy_burnin(d, group, variant, monkeypatch):
    # adapthist > 1 (MCMC_adapt.F90:116-136) and greedy burn-in (:83-101) on the warp-per-chain kernel: the tick kernel
    # works from the rows kept in the chain's row buffer
    nml = {"ap": dict(nsimu=700, adaptint=50, adapthist=80, drscale=2.0, initcmatn=1, updatesigma=0),
           "ap_short": dict(nsimu=700, adaptint=40, adapthist=30, drscale=0.0, initcmatn=1, updatesigma=1, N0=4.0, S02=1.0),
           "ap_burnin": dict(nsimu=700, adaptint=50, adapthist=60, burnintime=150, doburnin=1, badaptint=30, scalelimit=0.3,
                             drscale=2.0, initcmatn=1, updatesigma=0),
           "greedy": dict(nsimu=700, adaptint=50, burnintime=300, doburnin=1, badaptint=25, greedy=1, scalelimit=0.05,
                          drscale=2.0, initcmatn=3 * d, updatesigma=0),
           "greedy_no_reset": dict(nsimu=700, adaptint=60, burnintime=280, doburnin=1, badaptint=35, greedy=1,
                                   scalelimit=0.05, drscale=2.0, initcmatn=3 * d, updatesigma=0),
           "greedy_only": dict(nsimu=400, doadapt=0, adaptint=50, burnintime=300, doburnin=1, badaptint=25, greedy=1,
                               scalelimit=0.05, drscale=0.0, initcmatn=3 * d, updatesigma=0)}[variant]
    if group:
        monkeypatch.setenv("MCMCB_K2_GROUP", "1")  # the opt-in group-of-warps kernel logs its rows the same way
    if variant == "ap_short" and d == 12:
        pytest.skip("a 30-step window holds fewer than d+1 distinct rows: singular covariance, parity undefined")
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    N = 4
    u = np.random.default_rng(7000 + d).random((N, (4 * d + 40) * nml["nsimu"]))
    par0, cmat0 = np.zeros(d), np.eye(d) * 0.5
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u, splits=[333, 1, nml["nsimu"] - 1 - 334])
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u, rtol=1e-9, at_tick=False)
    s.close()


def test_early_rejection_by_the_prior_alone():
    # a Gaussian prior tighter than the likelihood: `sspri2 >= sscrit` fires (erstayed, MCMC_run_er.F90:62-67) and the
    # model is then not consulted; mcmcb_fetch("erstayed") on the warp-per-chain kernel
    d = 6
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(method="er", nsimu=400, adaptint=100, initcmatn=3, updatesigma=0)
    prior = (np.full(d, 0.3), np.full(d, 0.2))
    N = 4
    u = np.random.default_rng(321).random((N, (4 * d + 40) * nml["nsimu"]))
    par0, cmat0 = np.zeros(d), np.eye(d) * 0.5
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u, prior=prior)
    er = s.fetch("erstayed")
    cnt = s.counters()
    for k in range(N):
        r = run_oracle(nml, k, O.MODEL_GAUSS, blob, par0, cmat0, u=u, prior=prior)
        assert (int(er[k]), int(cnt["stayed"][k]), int(cnt["ndrawn"][k])) == (r["erstayed"], r["stayed"], r["ndrawn"])
        np.testing.assert_allclose(s.fetch("par")[k], r["par"], rtol=1e-9, atol=1e-12)
    assert int(er.sum()) > 0
    s.close()


def test_c2_shape_philox():
    # BASELINE config C2 shape: d=100 correlated Gaussian, DRAM, per-chain private factor in HBM
    d = 100
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=401, adaptint=200, drscale=2.0, initcmatn=1, updatesigma=0)
    N = 64
    par0 = np.zeros(d)
    cmat0 = 0.01 * np.eye(d)
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, seed=12345)
    assert s.info()["kernel"] == 2
    compare(s, nml, [0, 63], O.MODEL_GAUSS, blob, par0, cmat0, seed=12345, rtol=1e-9)
    s.close()


def test_ram_banana_parity():
    d = 6
    blob = mb.models.blob_banana(d, 0.03)
    nml = dict(method="ram", nsimu=500, updatesigma=0, alphatarget=0.234, nuparam=0.7)
    N = 4
    u = np.random.default_rng(3).random((N, 30 * 500))
    par0, cmat0 = np.zeros(d), np.eye(d)
    s = run_gpu(nml, N, "banana", blob, par0, cmat0, u=u)
    compare(s, nml, range(N), O.MODEL_BANANA, blob, par0, cmat0, u=u, rtol=1e-9)
    s.close()


def test_ram_c4_shape_philox():
    d = 50
    blob = mb.models.blob_banana(d, 0.03)
    nml = dict(method="ram", nsimu=301, updatesigma=0, alphatarget=0.234, nuparam=0.7)
    par0, cmat0 = np.zeros(d), np.eye(d)
    s = run_gpu(nml, 40, "banana", blob, par0, cmat0, seed=5)
    compare(s, nml, [0, 39], O.MODEL_BANANA, blob, par0, cmat0, seed=5, rtol=1e-8)
    s.close()


def test_hier_model_dram():
    rng = np.random.default_rng(1)
    G, J = 6, 5
    y = rng.normal(size=(G, 1)) * 2 + rng.normal(size=(G, J))
    blob = mb.models.blob_hier(y)
    d = G + 2
    nml = dict(nsimu=400, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=0)
    par0 = np.r_[y.mean(1), 0.0, 0.0]
    cmat0 = 0.05 * np.eye(d)
    s = run_gpu(nml, 3, "hier", blob, par0, cmat0, seed=2)
    compare(s, nml, range(3), O.MODEL_HIER, blob, par0, cmat0, seed=2)
    s.close()


def test_resume_and_sharding_bit_identical():
    d = 20
    mu, lam, Sig = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=301, adaptint=60, drscale=2.0, initcmatn=1, updatesigma=1)
    par0, cmat0 = np.zeros(d), 0.1 * np.eye(d)
    a = run_gpu(nml, 16, "gauss", blob, par0, cmat0, seed=4)
    b = run_gpu(nml, 16, "gauss", blob, par0, cmat0, seed=4, splits=[0, 1, 58, 1, 120, 120])
    c = run_gpu(nml, 8, "gauss", blob, par0, cmat0, seed=4, chain_offset=8)
    for what in ("par", "ss", "sigma2", "R", "cmat", "mean", "counters"):
        assert np.array_equal(a.fetch(what), b.fetch(what)), what
    assert np.array_equal(a.fetch("par")[8:], c.fetch("par"))
    a.close(); b.close(); c.close()


def test_posterior_statistics_match_target():
    # statistical check (3): pooled over chains, mean ~ mu and covariance ~ Sigma (sigma2 fixed at 1)
    d = 8
    mu, lam, Sig = gauss_target(d, rho=0.5)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=3001, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=0)
    N = 512
    s = run_gpu(nml, N, "gauss", blob, np.zeros(d), 0.1 * np.eye(d), seed=77)
    s.run(0)
    par = s.fetch("par")
    se = np.sqrt(np.diag(Sig) / N)
    assert (np.abs(par.mean(0)) < 5 * se).all()
    emp = np.cov(par.T)
    assert (np.abs(emp - Sig) < 0.35 * np.sqrt(np.outer(np.diag(Sig), np.diag(Sig)))).all()
    acc = 1 - s.counters()["stayed"].mean() / 3000
    assert 0.2 < acc < 0.9
    s.close()

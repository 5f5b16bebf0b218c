"""GPU parity of the SVD-factor paths (SCAM sampler, usesvd DRAM/AM) against the CPU oracle.

The device forms a SCAM component proposal as theta + delta*U(:,j) where the reference (and the
oracle) computes U (U' theta + delta e_j) (MCMC_run_scam.F90:122-138): rounding-level
differences, so accept counts / chain indices / draw counts must still be bit-exact under the
same draws while values are compared at 1e-9.  Targets are chosen with well separated
covariance eigenvalues: singular vectors of a degenerate covariance are not unique."""
import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def aniso_gauss(d, seed=0):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    ev = np.geomspace(0.05, 4.0, d)
    Sig = (Q * ev) @ Q.T
    Sig = 0.5 * (Sig + Sig.T)
    lam = np.linalg.inv(Sig)
    return np.zeros(d), 0.5 * (lam + lam.T), Sig


def run_gpu(nml, N, model, blob, par0, cmat0, u=None, seed=0, splits=None, chain_offset=0, sigma2=(1.0,), nobs=(1,)):
    cfg = mb.default_config(nchains=N, seed=seed, store_chains=-1, model=model, chain_offset=chain_offset,
                            rng_mode=mb.RNG_INJECTED if u is not None else mb.RNG_PHILOX, **nml)
    s = mb.Sampler(cfg)
    s.set_data(blob)
    s.set_initial(par0, cmat0, list(sigma2), list(nobs))
    if u is not None:
        s.inject_uniforms(u)
    for n in (splits or [nml["nsimu"] - 1]):
        s.run(n)
    return s


def run_oracle(nml, k, model_id, blob, par0, cmat0, u=None, seed=0, chain_offset=0, sigma2=(1.0,), nobs=(1,)):
    ch = O.Chain(O.make_cfg(**nml), model_id, blob, par0, cmat0, list(sigma2), list(nobs))
    if u is not None:
        ch.inject(u[k])
    else:
        ch.philox(seed, chain_offset + k)
    ch.run()
    return ch.results()


def compare(s, nml, ks, model_id, blob, par0, cmat0, u=None, seed=0, rtol=1e-9, sigma2=(1.0,), nobs=(1,), scam=True):
    cnt = s.counters()
    par, ss, s2 = s.fetch("par"), s.fetch("ss"), s.fetch("sigma2")
    R, cm, mean, wsum = s.fetch("R"), s.fetch("cmat"), s.fetch("mean"), s.fetch("wsum")
    q = s.fetch("qcovstd")
    for k in ks:
        r = run_oracle(nml, k, model_id, blob, par0, cmat0, u=u, seed=seed, sigma2=sigma2, nobs=nobs)
        for key in ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "status", "ndrawn"):
            assert cnt[key][k] == r[key], (k, key, cnt[key][k], r[key])
        g = s.fetch_chain(k)
        assert np.array_equal(g["chain"][:, -1], r["chain"][:, -1])
        scale = np.abs(r["chain"][:, :-1]).max()
        np.testing.assert_allclose(g["chain"][:, :-1], r["chain"][:, :-1], rtol=rtol, atol=rtol * scale)
        np.testing.assert_allclose(g["sschain"][:, :-1], r["sschain"][:, :-1], rtol=1e-8)
        np.testing.assert_allclose(g["s2chain"][:r["simuind"]], r["s2chain"][:r["simuind"]], rtol=1e-8)
        np.testing.assert_allclose(par[k], r["par"], rtol=rtol, atol=rtol * scale)
        np.testing.assert_allclose(ss[k], np.atleast_1d(r["sschain"][-1, 0]), rtol=1e-8)
        # factor: same eigenvectors (sign convention included) and scales
        np.testing.assert_allclose(R[k], r["R"], rtol=0, atol=1e-7 * np.abs(r["R"]).max())
        if scam:
            np.testing.assert_allclose(q[k], r["qcovstd"], rtol=1e-8)
        ns, ai = nml["nsimu"], nml.get("adaptint", 100)
        if ns % ai == 0:
            assert wsum[k, 0] == r["wsum"]
            np.testing.assert_allclose(mean[k], r["mean"], rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(cm[k], r["cmat"], rtol=0, atol=1e-8 * np.abs(r["cmat"]).max())


@pytest.mark.parametrize("d", [3, 7])
@pytest.mark.parametrize("sig", [0, 1])
def test_scam_gauss_injected(d, sig):
    mu, lam, Sig = aniso_gauss(d, seed=d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(method="scam", nsimu=600, adaptint=150, initcmatn=2, updatesigma=sig, N0=3.0, S02=1.0)
    N = 4
    u = np.random.default_rng(10 + d).random((N, (4 * d + 30) * nml["nsimu"]))
    par0 = np.zeros(d)
    cmat0 = np.diag(np.linspace(0.2, 0.6, d))  # distinct: the initial eigen-basis is unambiguous
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u, sigma2=(1.0,), nobs=(d,))
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u, sigma2=(1.0,), nobs=(d,))
    s.close()


@pytest.mark.parametrize("mode", ["scam", "usesvd"])
def test_ap_window_with_svd_factors(mode):
    # adapthist > 1 (MCMC_adapt.F90:116-136) feeding the SVD square root: SCAM, and AM with condmax > 0
    d = 4
    mu, lam, Sig = aniso_gauss(d, seed=11)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=640, adaptint=80, adapthist=120, initcmatn=2, updatesigma=0)
    nml.update(dict(method="scam") if mode == "scam" else dict(condmax=1e12, drscale=0.0))
    N = 3
    u = np.random.default_rng(50).random((N, (4 * d + 30) * nml["nsimu"]))
    par0, cmat0 = np.zeros(d), np.diag(np.linspace(0.2, 0.6, d))
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u, nobs=(d,), splits=[333, 306])
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u, nobs=(d,), scam=(mode == "scam"))
    s.close()


def test_scam_hier_philox_with_bounds_free_model():
    rng = np.random.default_rng(1)
    G, J = 6, 5
    y = rng.normal(size=(G, 1)) * 2 + rng.normal(size=(G, J))
    blob = mb.models.blob_hier(y)
    d = G + 2
    nml = dict(method="scam", nsimu=500, adaptint=125, initcmatn=1, updatesigma=0)
    par0 = np.r_[y.mean(1), 0.0, 0.0]
    cmat0 = np.diag(np.linspace(0.03, 0.08, d))
    s = run_gpu(nml, 3, "hier", blob, par0, cmat0, seed=2)
    compare(s, nml, range(3), O.MODEL_HIER, blob, par0, cmat0, seed=2)
    s.close()


def test_scam_config_rules():
    # method='scam' forces condmax=1e15, doburnin=0, DR off (mcmcinit.F90:324-333)
    d = 3
    mu, lam, Sig = aniso_gauss(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(method="scam", nsimu=120, adaptint=60, initcmatn=1, updatesigma=0, drscale=3.0, doburnin=1,
               burnintime=50)
    s = run_gpu(nml, 2, "gauss", blob, np.zeros(d), np.diag([0.2, 0.3, 0.5]), seed=9)
    compare(s, nml, range(2), O.MODEL_GAUSS, blob, np.zeros(d), np.diag([0.2, 0.3, 0.5]), seed=9)
    assert (s.counters()["drtries"] == 0).all()
    s.close()


def test_scam_resume_bit_identical():
    d = 6
    mu, lam, Sig = aniso_gauss(d, seed=3)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(method="scam", nsimu=301, adaptint=60, initcmatn=1, updatesigma=1)
    par0, cmat0 = np.zeros(d), np.diag(np.linspace(0.1, 0.4, d))
    a = run_gpu(nml, 8, "gauss", blob, par0, cmat0, seed=4)
    b = run_gpu(nml, 8, "gauss", blob, par0, cmat0, seed=4, splits=[0, 1, 58, 1, 120, 120])
    c = run_gpu(nml, 4, "gauss", blob, par0, cmat0, seed=4, chain_offset=4)
    for what in ("par", "ss", "sigma2", "R", "qcovstd", "cmat", "mean", "counters"):
        assert np.array_equal(a.fetch(what), b.fetch(what)), what
    assert np.array_equal(a.fetch("par")[4:], c.fetch("par"))
    a.close(); b.close(); c.close()


def test_scam_posterior_statistics():
    d = 5
    mu, lam, Sig = aniso_gauss(d, seed=5)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(method="scam", nsimu=2001, adaptint=100, initcmatn=1, updatesigma=0)
    N = 512
    s = run_gpu(nml, N, "gauss", blob, np.zeros(d), 0.1 * np.eye(d), seed=77)
    par = s.fetch("par")
    se = np.sqrt(np.diag(Sig) / N)
    assert (np.abs(par.mean(0)) < 5 * se).all()
    emp = np.cov(par.T)
    assert (np.abs(emp - Sig) < 0.35 * np.sqrt(np.outer(np.diag(Sig), np.diag(Sig)))).all()
    assert (s.counters()["status"] == 0).all()
    s.close()


@pytest.mark.parametrize("d", [4, 9])
def test_usesvd_am_injected(d):
    # condmax > 0: covtor_svd factor, proposal theta + R z (matutils.F90:378-453, MCMC_DRAM.F90:27)
    mu, lam, Sig = aniso_gauss(d, seed=20 + d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=600, adaptint=100, drscale=0.0, initcmatn=2, updatesigma=0, condmax=1e10)
    N = 4
    u = np.random.default_rng(d).random((N, (4 * d + 30) * nml["nsimu"]))
    par0, cmat0 = np.zeros(d), np.diag(np.linspace(0.2, 0.6, d))
    s = run_gpu(nml, N, "gauss", blob, par0, cmat0, u=u)
    compare(s, nml, range(N), O.MODEL_GAUSS, blob, par0, cmat0, u=u, scam=False)
    s.close()


def test_usesvd_condmax_floor_rewrites_cmat():
    # a tiny condmax floors the small singular values and replaces cmat by R0 R0' (MCMC_adapt.F90:205-208)
    d = 5
    mu, lam, Sig = aniso_gauss(d, seed=31)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=400, adaptint=100, drscale=0.0, initcmatn=2, updatesigma=0, condmax=3.0)
    par0, cmat0 = np.zeros(d), np.diag(np.linspace(0.2, 0.6, d))
    s = run_gpu(nml, 3, "gauss", blob, par0, cmat0, seed=8)
    compare(s, nml, range(3), O.MODEL_GAUSS, blob, par0, cmat0, seed=8, scam=False)
    s.close()


def test_unsupported_combinations_are_refused():
    d = 4
    mu, lam, Sig = aniso_gauss(d)
    blob = mb.models.blob_gauss(mu, lam)
    for nml in (dict(nsimu=10, condmax=1e8, drscale=2.0), dict(nsimu=10, method="ram", condmax=1e8)):
        s = mb.Sampler(mb.default_config(nchains=2, model="gauss", **nml))
        s.set_data(blob)
        with pytest.raises(mb.MCMCBError):
            s.set_initial(np.zeros(d), np.eye(d), [1.0], [1])
        s.close()

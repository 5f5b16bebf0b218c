"""GPU parity of the rows SURVEY.md 8f ranks N4, on the register kernel (K1), through the C ABI:
the early-rejection sampler (MCMC_run_er.F90:12-107, MCMC_sscrit MCMC_DRAM.F90:124-135), AP windowed
adaptation (adapthist > 1, MCMC_adapt.F90:116-136) and greedy burn-in (MCMC_adapt.F90:83-101).
Same bars as test_k1_parity.py: integers bit-exact, theta / ss within the stated relative tolerance."""
import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O
from tests import cases
from tests.test_k1_parity import BLOB11, _compare, _gpu_run, _oracle_run

pytestmark = pytest.mark.gpu

ER = dict(method="er", nsimu=801, adaptint=100, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0)


def _erstayed(nml, k, blob, par0, u=None, seed=0, chain_offset=0, prior=None, **kw):
    ch = O.Chain(O.make_cfg(**nml), O.MODEL_EXPREG, blob, par0, kw.get("cmat0", cases.CMAT0),
                 kw.get("sigma2", cases.SIGMA2), kw.get("nobs", cases.NOBS), prior=prior)
    if u is not None:
        ch.inject(u[k])
    else:
        ch.philox(seed, chain_offset + k)
    ch.run()
    return ch.counters()["erstayed"]


@pytest.mark.parametrize("lanes", [1, 4, 32])
def test_er_injected_draws(lanes):
    N = 5
    u = np.random.default_rng(77).random((N, 12 * ER["nsimu"]))
    s = _gpu_run(ER, N, BLOB11, cases.PAR0, u=u, lanes=lanes)
    _compare(s, ER, N, BLOB11, cases.PAR0, u=u, RTOL=1e-11)
    assert (s.fetch("erstayed") == 0).all()  # flat prior: nothing is rejected by the prior alone
    s.close()


def test_er_prior_rejections_and_dr_switched_off():
    # a tight prior makes `sspri2 >= sscrit` fire (erstayed, MCMC_run_er.F90:62-67); drscale > 0 is ignored
    # ("no dr with er", MCMC_run_er.F90:24-27)
    nml = dict(ER, drscale=2.0, nsimu=601)
    prior = (np.array([9.8, 0.1]), np.array([0.05, 0.002]))
    N = 4
    u = np.random.default_rng(5).random((N, 12 * nml["nsimu"]))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u, prior=prior)
    _compare(s, nml, N, BLOB11, cases.PAR0, u=u, prior=prior, check_factors=False, RTOL=1e-8)  # tight posterior: rounding of the adapted factor is amplified
    er = s.fetch("erstayed")
    ref = [_erstayed(nml, k, BLOB11, cases.PAR0, u=u, prior=prior) for k in range(N)]
    assert list(er) == ref and min(ref) > 0
    cnt = s.counters()
    assert (cnt["drtries"] == 0).all()
    s.close()


@pytest.mark.parametrize("early_exit", [0, 1])
def test_er_early_exit_kernel_equals_full_sum(early_exit, monkeypatch):
    # thread per chain on 10^4 data in shared memory: the warp-vote early exit of ExpReg::ssfunction_er leaves
    # every chain identical to the oracle's full-sum default (ssfunction_er0.f90); resumed across launches
    # (early_exit = 0: the default for 'er', four chains per thread through ssfunction_batch)
    monkeypatch.setenv("MCMCB_ER_EXIT", str(early_exit))
    x, y = cases.synth_expreg(10000)
    blob = mb.models.blob_expreg(x, y)
    nml = dict(method="er", nsimu=241, adaptint=60, initcmatn=1, updatesigma=1, N0=1.0, S02=0.5)
    N = 70
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(3).normal(size=(N, 2)))
    cm0 = cases.CMAT0 * 11.0 / 10000
    s = _gpu_run(nml, N, blob, par0, seed=11, lanes=1, cmat0=cm0, nobs=[10000], splits=[100, 1, 139])
    import os
    want = 1 if early_exit else int(os.environ.get("MCMCB_K1_BATCH", "4"))  # (a tuning override may be in force)
    assert s.info()["lanes_per_chain"] == 1 and s.info()["chains_per_thread"] == want
    _compare(s, nml, 6, blob, par0, seed=11, cmat0=cm0, nobs=[10000], RTOL=1e-10)
    s.close()


AP = dict(cases.NML_DRAM, nsimu=901, adaptint=50, adapthist=80)


@pytest.mark.parametrize("name", ["ap", "ap_short_window", "ap_burnin", "ap_am"])
def test_ap_window(name):
    nml = {
        "ap": AP,
        "ap_short_window": dict(AP, adapthist=25, adaptint=30),         # ring wraps many times
        "ap_burnin": dict(AP, doburnin=1, burnintime=200, badaptint=40, scalelimit=0.3),
        "ap_am": dict(AP, drscale=0.0, updatesigma=0, adapthist=120, adaptend=700),
    }[name]
    N = 5
    u = np.random.default_rng(21).random((N, 40 * nml["nsimu"]))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u, splits=[333, 1, nml["nsimu"] - 1 - 334])
    _compare(s, nml, N, BLOB11, cases.PAR0, u=u, check_factors=False, RTOL=1e-9)
    r = _oracle_run(nml, 0, BLOB11, cases.PAR0, u=u)
    iu = np.triu_indices(2)
    np.testing.assert_allclose(s.fetch("R")[0][iu], r["R"][iu], rtol=1e-8)
    s.close()


def test_ap_accumulators_at_the_tick():
    # stop exactly at an AP tick: chaincmat / chainmean / chainwsum are the window's batch moments
    nml = dict(AP, nsimu=600)
    N = 3
    u = np.random.default_rng(22).random((N, 40 * nml["nsimu"]))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u)
    iu = np.triu_indices(2)
    for k in range(N):
        r = _oracle_run(nml, k, BLOB11, cases.PAR0, u=u)
        assert s.fetch("wsum")[k, 0] == r["wsum"] == nml["adapthist"]
        np.testing.assert_allclose(s.fetch("mean")[k], r["mean"], rtol=1e-11)
        np.testing.assert_allclose(s.fetch("cmat")[k][iu], r["cmat"][iu], rtol=1e-8)
    s.close()


GREEDY = dict(nsimu=901, adaptint=50, burnintime=300, doburnin=1, badaptint=25, greedy=1, scalelimit=0.05,
              drscale=2.0, initcmatn=3, updatesigma=1, N0=1.0, S02=0.0)
# (initcmatn = 1 gives cmat0 zero weight in the covmat recursion: with the two or three rows a chain has at its
# first greedy tick the covariance is then singular and whether Cholesky "succeeds" is rounding noise)


@pytest.mark.parametrize("name", ["reset", "no_reset", "am", "burnin_only", "scaling"])
def test_greedy_burnin(name):
    nml = {
        "reset": GREEDY,                                                  # 300+50 is a tick: accumulators restart
        "no_reset": dict(GREEDY, burnintime=280, adaptint=60, badaptint=35),  # 340 is no tick: AM carries on from the greedy covariance
        "am": dict(GREEDY, drscale=0.0, updatesigma=0, initcmatn=5),
        "burnin_only": dict(GREEDY, doadapt=0, nsimu=401),
        "scaling": dict(GREEDY, scalelimit=0.4),                          # scale ticks and greedy ticks interleave
    }[name]
    N = 5
    u = np.random.default_rng(31).random((N, 40 * nml["nsimu"]))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u, splits=[150, 1, nml["nsimu"] - 1 - 151])
    _compare(s, nml, N, BLOB11, cases.PAR0, u=u, check_factors=False, RTOL=1e-9)
    iu = np.triu_indices(2)
    for k in range(N):
        r = _oracle_run(nml, k, BLOB11, cases.PAR0, u=u)
        np.testing.assert_allclose(s.fetch("R")[k][iu], r["R"][iu], rtol=1e-8)
    s.close()


def test_greedy_factor_right_after_a_greedy_tick():
    nml = dict(GREEDY, nsimu=250)  # last tick (250 < burnintime) is a burn-in tick
    N = 4
    u = np.random.default_rng(32).random((N, 40 * nml["nsimu"]))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u)
    iu = np.triu_indices(2)
    for k in range(N):
        r = _oracle_run(nml, k, BLOB11, cases.PAR0, u=u)
        for f in ("R", "R2", "iC"):
            np.testing.assert_allclose(s.fetch(f)[k][iu], r[f][iu], rtol=1e-9, err_msg=f)
    s.close()


def test_svd_factor_refuses_greedy_burnin():
    # greedy burn-in exists for the Cholesky-factor samplers (K1, K2: tests/test_k2_parity.py), not with condmax > 0
    cfg = mb.default_config(nchains=4, model="hier", nsimu=100, adaptint=20, condmax=1e10, greedy=1, doburnin=1,
                            burnintime=50)
    with pytest.raises(mb.MCMCBError) as e:
        mb.Sampler(cfg)
    assert "EUNSUPPORTED" in str(e.value)

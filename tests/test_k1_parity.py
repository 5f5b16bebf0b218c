"""GPU parity of the small-npar register kernel (K1) against the CPU oracle, through the
C ABI.  Bars (BASELINE.json north_star): (1) chain indices / accept counts bit-exact under
injected draws; (2) theta and ss within 1e-12 relative per step."""
import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _gpu_run(nml, N, blob, par0, u=None, seed=0, lanes=0, cmat0=cases.CMAT0, sigma2=cases.SIGMA2, nobs=cases.NOBS,
             steps=None, chain_offset=0, splits=None, model="expreg", prior=None):
    cfg = mb.default_config(nchains=N, seed=seed, store_chains=-1, model=model, lanes_per_chain=lanes,
                            rng_mode=mb.RNG_INJECTED if u is not None else mb.RNG_PHILOX,
                            chain_offset=chain_offset, **nml)
    s = mb.Sampler(cfg)
    s.set_data(blob)
    if prior is not None:
        s.set_priors(*prior)
    s.set_initial(par0, cmat0, sigma2, nobs)
    if u is not None:
        s.inject_uniforms(u)
    total = nml["nsimu"] - 1 if steps is None else steps
    for n in (splits or [total]):
        s.run(n)
    return s


def _oracle_run(nml, k, blob, par0, u=None, seed=0, model=O.MODEL_EXPREG, cmat0=cases.CMAT0, sigma2=cases.SIGMA2,
                nobs=cases.NOBS, chain_offset=0, prior=None):
    ch = O.Chain(O.make_cfg(**nml), model, blob, par0, cmat0, sigma2, nobs, prior=prior)
    if u is not None:
        ch.inject(u[k])
    else:
        ch.philox(seed, chain_offset + k)
    ch.run()
    return ch.results()


def _compare(s, nml, N, blob, par0, u=None, seed=0, chain_offset=0, check_factors=True, prior=None,
             cmat0=cases.CMAT0, sigma2=cases.SIGMA2, nobs=cases.NOBS, model=O.MODEL_EXPREG, RTOL=RTOL):
    cnt = s.counters()
    par, ss, s2 = s.fetch("par"), s.fetch("ss"), s.fetch("sigma2")
    mean, cm, R, wsum = s.fetch("mean"), s.fetch("cmat"), s.fetch("R"), s.fetch("wsum")
    R2, iC = s.fetch("R2"), s.fetch("iC")
    iu = np.triu_indices(par.shape[1])
    for k in range(N):
        p0 = par0 if np.ndim(par0) == 1 else par0[k]
        r = _oracle_run(nml, k, blob, p0, u=u, seed=seed, chain_offset=chain_offset, prior=prior, cmat0=cmat0,
                        sigma2=sigma2, nobs=nobs, model=model)
        for key in ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "status", "ndrawn"):
            assert cnt[key][k] == r[key], (k, key)          # bar (1): bit-exact integers
        g = s.fetch_chain(k)
        assert g["nrows"] == r["chainind"]
        assert np.array_equal(g["chain"][:, -1], r["chain"][:, -1])      # run-length counts
        assert np.array_equal(g["sschain"][:, -1], r["sschain"][:, -1])
        np.testing.assert_allclose(g["chain"][:, :-1], r["chain"][:, :-1], rtol=RTOL, atol=0)     # bar (2)
        np.testing.assert_allclose(g["sschain"][:, :-1], r["sschain"][:, :-1], rtol=RTOL, atol=0)
        n2 = r["simuind"]
        np.testing.assert_allclose(g["s2chain"][:n2], r["s2chain"][:n2], rtol=RTOL, atol=0)
        np.testing.assert_allclose(par[k], r["par"], rtol=RTOL)
        np.testing.assert_allclose(s2[k], r["sigma2"], rtol=RTOL)
        if check_factors:
            np.testing.assert_allclose(R[k][iu], r["R"][iu], rtol=max(1e-10, 10 * RTOL))
            if nml.get("drscale", 0) > 0 and nml.get("method", "dram") == "dram":
                np.testing.assert_allclose(R2[k][iu], r["R2"][iu], rtol=1e-10)
                np.testing.assert_allclose(iC[k][iu], r["iC"][iu], rtol=1e-9)
            ns, ai = nml["nsimu"], nml.get("adaptint", 100)
            at_tick = (nml.get("method", "dram") != "ram" and nml.get("doadapt", 1) and ns % ai == 0
                       and ns >= nml.get("burnintime", 0) + ai and not (0 < nml.get("adaptend", 0) < ns))
            if at_tick:
                # the streaming accumulators equal the reference's chaincmat/chainmean at adaptation ticks
                assert wsum[k, 0] == r["wsum"]
                np.testing.assert_allclose(mean[k], r["mean"], rtol=1e-11)
                np.testing.assert_allclose(cm[k][iu], r["cmat"][iu], rtol=1e-9)


BLOB11 = mb.models.blob_expreg(cases.DATA_X, cases.DATA_Y)


@pytest.mark.parametrize("nml_name", ["shipped", "dram", "dram_burnin", "am_nosigma", "initcmatn0"])
def test_injected_draws_bit_exact_counts(nml_name):
    nml = {
        "shipped": cases.NML_SHIPPED,
        "dram": dict(cases.NML_DRAM, nsimu=1000),
        "dram_burnin": dict(nsimu=1200, adaptint=100, burnintime=300, doburnin=1, badaptint=50, drscale=3.0,
                            initcmatn=2, scalelimit=0.2, updatesigma=1),
        "am_nosigma": dict(nsimu=800, adaptint=80, drscale=0.0, initcmatn=1, updatesigma=0, adaptend=600),
        "initcmatn0": dict(nsimu=600, adaptint=100, drscale=2.0, initcmatn=0, updatesigma=1),
    }[nml_name]
    N = 6
    u = np.random.default_rng(1234).random((N, 40 * nml["nsimu"]))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u)
    factors = nml_name != "initcmatn0"
    _compare(s, nml, N, BLOB11, cases.PAR0, u=u, check_factors=factors)
    if not factors:  # streaming start-up differs from the two-pass batch formula by rounding only
        r = _oracle_run(nml, 0, BLOB11, cases.PAR0, u=u)
        np.testing.assert_allclose(s.fetch("cmat")[0][np.triu_indices(2)], r["cmat"][np.triu_indices(2)], rtol=1e-8)
    s.close()


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
def test_philox_parity_all_lane_groupings(lanes):
    nml = dict(cases.NML_DRAM, nsimu=501)
    N = 37  # ragged: not a multiple of 32/lanes
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(7).normal(size=(N, 2)))
    s = _gpu_run(nml, N, BLOB11, par0, seed=99, lanes=lanes, chain_offset=1000)
    assert s.info()["lanes_per_chain"] == lanes
    _compare(s, nml, 5, BLOB11, par0, seed=99, chain_offset=1000)
    s.close()


def test_large_ndata_in_shared_memory():
    x, y = cases.synth_expreg(10000)
    blob = mb.models.blob_expreg(x, y)
    nml = dict(nsimu=101, adaptint=50, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.5)
    cm0 = cases.CMAT0 * (11.0 / 10000)
    N = 40
    for lanes in (1, 32):
        s = _gpu_run(nml, N, blob, cases.PAR0, seed=5, lanes=lanes, cmat0=cm0, sigma2=[0.5], nobs=[10000])
        assert s.info()["smem_bytes"] >= 160000
        _compare(s, nml, 3, blob, cases.PAR0, seed=5, cmat0=cm0, sigma2=[0.5], nobs=[10000])
        s.close()


def test_odd_ndata_and_gaussian_prior():
    x, y = cases.synth_expreg(37, seed=3)
    blob = mb.models.blob_expreg(x, y)
    prior = (np.array([9.0, 0.2]), np.array([2.0, 0.0]))  # second component disabled (sig <= 0)
    nml = dict(nsimu=400, adaptint=50, drscale=2.0, initcmatn=1)
    s = _gpu_run(nml, 4, blob, cases.PAR0, seed=2, lanes=4, nobs=[37], prior=prior)
    _compare(s, nml, 4, blob, cases.PAR0, seed=2, nobs=[37], prior=prior)
    s.close()


def test_resume_is_bit_identical():
    nml = dict(cases.NML_DRAM, nsimu=401)
    a = _gpu_run(nml, 16, BLOB11, cases.PAR0, seed=3)
    b = _gpu_run(nml, 16, BLOB11, cases.PAR0, seed=3, splits=[1, 99, 150, 150])
    for what in ("par", "ss", "sigma2", "R", "cmat", "mean", "counters"):
        assert np.array_equal(a.fetch(what), b.fetch(what)), what
    assert np.array_equal(a.fetch_chain(3)["chain"], b.fetch_chain(3)["chain"])
    a.close(); b.close()


def test_results_do_not_depend_on_sharding():
    # chain_offset shards one logical population over handles: same global ids -> same chains
    nml = dict(cases.NML_DRAM, nsimu=201)
    whole = _gpu_run(nml, 64, BLOB11, cases.PAR0, seed=8)
    part = _gpu_run(nml, 32, BLOB11, cases.PAR0, seed=8, chain_offset=32)
    assert np.array_equal(whole.fetch("par")[32:], part.fetch("par"))
    assert np.array_equal(whole.fetch("counters")[32:, :7], part.fetch("counters")[:, :7])
    whole.close(); part.close()


def test_ram_parity():
    nml = dict(method="ram", nsimu=600, updatesigma=1, alphatarget=0.234, nuparam=0.7)
    N = 5
    u = np.random.default_rng(77).random((N, 30 * 600))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u)
    # counts stay bit-exact; values get 1e-9: dchdd's downdate is ill-conditioned when |a| -> 1
    # (error amplification 1/sqrt(1-|a|^2), dchdd.f:149-154) and this 11-point target drives it there
    # (several downdates fail outright, status bit 2), so last-bit libm differences (pow, exp) grow
    _compare(s, nml, N, BLOB11, cases.PAR0, u=u, RTOL=1e-9)
    s.close()


@pytest.mark.parametrize("cm0", [np.diag([4.0, 0.02]), np.diag([400.0, 4.0])], ids=["wide", "extreme"])
def test_out_of_bounds_and_extreme_proposals(cm0):
    # wide proposals: many violate theta>0 (Q3) or underflow alpha12 to exactly 0 (Q1); the extreme
    # case also drives alpha13 = exp(l2+q1)*... into the subnormal range, where the reference has no
    # clamp (MCMC_DRAM.F90:184) and "draw a uniform iff alpha>0" hinges on the last subnormal bit
    # (exp_subnormal_safe in csrc/common.cuh).  Adaptation is off: a chain that barely moves feeds a
    # rank-deficient covariance to the Cholesky, whose success is then decided by rounding noise.
    nml = dict(nsimu=500, doadapt=0, drscale=2.0, updatesigma=1)
    N = 4
    u = np.random.default_rng(5).random((N, 40 * 500))
    s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u, cmat0=cm0)
    _compare(s, nml, N, BLOB11, cases.PAR0, u=u, cmat0=cm0)
    assert s.counters()["bndstayed"].sum() > 0
    s.close()


def test_injected_stream_exhaustion_is_flagged():
    nml = dict(nsimu=200, doadapt=0, updatesigma=0)
    u = np.random.default_rng(0).random((2, 50))
    s = _gpu_run(nml, 2, BLOB11, cases.PAR0, u=u)
    assert (s.counters()["status"] & 4).all()
    s.close()


def test_unsupported_configurations_fail_loudly():
    # (AP windows and greedy burn-in are built in this kernel: tests/test_k1_er_ap_greedy.py)
    # (SVD factors -- method='scam', condmax > 0 -- run on the warp-per-chain kernels; forcing the register kernel fails)
    for kw in (dict(condmax=1e10, kernel=1), dict(method="scam", kernel=1), dict(adapthist=50, pool_adapt=1)):
        with pytest.raises(mb.MCMCBError):
            mb.Sampler(mb.default_config(nsimu=10, nchains=2, **kw))
    with pytest.raises(mb.MCMCBError):
        mb.Sampler(mb.default_config(nsimu=10, nchains=2, model="no-such-model"))


def test_streamed_dumps():
    nml = dict(cases.NML_DRAM, nsimu=301)
    cfg = mb.default_config(nchains=100, seed=1, dump_stride=100, **nml)
    s = mb.Sampler(cfg)
    s.set_data(BLOB11)
    s.set_initial(cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS)
    s.run(300)
    got = []
    while True:
        d = s.dump_pop()
        if d is None:
            break
        got.append(d)
    assert [g[0] for g in got] == [101, 201, 301]
    assert np.array_equal(got[-1][1], s.fetch("par"))
    s.close()


def test_full_size_properties_one_million_chains():
    # BASELINE config C3 population size at a short length: size-independent invariants
    N = 1 << 20
    x, y = cases.synth_expreg(10000)
    blob = mb.models.blob_expreg(x, y)
    nml = dict(nsimu=5, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.5)
    rng = np.random.default_rng(0)
    par0 = cases.PAR0 * (1 + 0.01 * rng.normal(size=(N, 2)))
    cfg = mb.default_config(nchains=N, seed=2024, store_chains=0, **nml)
    s = mb.Sampler(cfg)
    s.set_data(blob)
    s.set_initial(par0, cases.CMAT0 * (11.0 / 10000), [0.5], [10000])
    s.run(4)
    c = s.counters()
    assert (c["simuind"] == 5).all()
    assert (c["status"] == 0).all()
    assert (c["chainind"] + c["stayed"] == 5).all()
    assert (c["draccepted"] <= c["drtries"]).all() and (c["drtries"] >= c["stayed"]).all()
    w = s.fetch("wsum")[:, 0]   # initcmatn + weight of the completed rows; the open row is still pending
    assert ((w >= 1) & (w <= 5)).all()
    par = s.fetch("par")
    assert np.isfinite(par).all() and (par > 0).all()
    # spot-check three chains against the oracle
    for k in (0, 12345, N - 1):
        r = _oracle_run(nml, k, blob, par0[k], seed=2024, cmat0=cases.CMAT0 * (11.0 / 10000), sigma2=[0.5],
                        nobs=[10000])
        assert c["stayed"][k] == r["stayed"]
        np.testing.assert_allclose(par[k], r["par"], rtol=RTOL)
    s.close()


@pytest.mark.parametrize("N", [1, 33, 400003])
def test_chains_per_thread_do_not_change_results(N, monkeypatch):
    # the register kernel runs 4 chains per thread where there are enough chains (tile plan of 4-, 2- and 1-sub-tile
    # tiles, csrc/launchers.cuh); a chain's result must not depend on how many neighbours share its thread.  400003
    # chains: 4-sub-tile tiles plus a ragged last round.
    nml = dict(cases.NML_DRAM, nsimu=100, adaptint=4, initcmatn=10)
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(17).normal(size=(N, 2)))
    out = {}
    for b in (1, 2, 4):
        monkeypatch.setenv("MCMCB_K1_BATCH", str(b))
        cfg = mb.default_config(nchains=N, seed=5, store_chains=0, model="expreg", lanes_per_chain=1, **nml)
        s = mb.Sampler(cfg)
        s.set_data(BLOB11)
        s.set_initial(par0, cases.CMAT0, cases.SIGMA2, cases.NOBS)
        s.run(3)
        s.run(6)
        assert s.info()["chains_per_thread"] == b
        out[b] = (s.fetch("par"), s.fetch("cmat"), s.fetch("sigma2"), s.fetch("counters"))
        s.close()
    for b in (2, 4):
        for x, y in zip(out[1], out[b]):
            assert np.array_equal(x, y)
    assert (out[1][3][:, 5] == 10).all() and (out[1][3][:, 6] == 0).all()  # simuind, status

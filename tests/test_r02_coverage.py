"""Round-2 GPU parity cases the round-1 suite did not reach (VERDICT "untested configs", NS1, N3, ADVICE):

* BASELINE C3 at its batched shape: n = 10^4 data in shared memory, FOUR chains per thread (the tile plan only hands
  out 4-chain tiles from ~3*10^5 chains up), 399 steps = 4 adaptation ticks with delayed rejection, vs the oracle;
* BASELINE C5's dimension: SCAM on the 200-parameter hierarchical model (CTA Jacobi at 200 x 200), vs the oracle;
* BASELINE C4's shape: pooled RAM at d = 50 on the banana target, vs the lock-step oracle;
* streamed dumps: EVERY popped snapshot (theta, ss, sigma2) vs the oracle's state at that step, K1 and K2;
* restart files: values of mcmccovf.dat / mcmcmean.dat vs the oracle, and a two-leg run through nmlffile + *f.dat
  (MCMC_aux.F90:46-83) whose second leg equals an oracle run started from the same files;
* one handle driving several GPUs (ngpus > 1): shards equal a single-GPU handle bit for bit; pooled ticks over NCCL;
* a burn-in long enough that the tick kernels' row buffer passes 48 KB (ADVICE: launch used to fail).
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O
from tests import cases
from tests.test_host_driver import HOST, NML_SHIPPED, write_testcase

pytestmark = pytest.mark.gpu
BLOB11 = mb.models.blob_expreg(cases.DATA_X, cases.DATA_Y)
CNT = ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "ndrawn")


def ndev():
    import torch
    return torch.cuda.device_count()


def gauss_target(d, rho=0.6):
    s = 1.0 + 2.0 * np.arange(d) / max(d - 1, 1)
    sig = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d))) * np.outer(s, s)
    lam = np.linalg.inv(sig)
    return np.zeros(d), 0.5 * (lam + lam.T)


def oracle_chain(nml, model_id, blob, par0, cmat0, sigma2, nobs, seed, chain_id):
    ch = O.Chain(O.make_cfg(**nml), model_id, blob, par0, cmat0, sigma2, nobs)
    ch.philox(seed, chain_id)
    return ch


# ------------------------------------------------------------------------------------------------ C3, 4 chains/thread
def test_c3_four_chains_per_thread_four_ticks_with_dr():
    # enough chains for full rounds of 4-chain tiles plus smaller tiles; simuind = 400 at the end = the 4th adaptation
    # tick, where the streaming accumulators equal the reference's chainmean / chaincmat
    N, steps = 148 * 512 * 4 + 4096, 399
    x, y = cases.synth_expreg(10000)
    blob = mb.models.blob_expreg(x, y)
    nml = dict(nsimu=steps + 1, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.5)
    rng = np.random.default_rng(11)
    par0 = cases.PAR0 * (1 + 0.01 * rng.normal(size=(N, 2)))
    cmat0 = cases.CMAT0 * (11.0 / 10000)
    s = mb.Sampler(mb.default_config(nchains=N, seed=2024, store_chains=0, **nml))
    s.set_data(blob)
    s.set_initial(par0, cmat0, [0.5], [10000])
    s.run(steps)
    assert s.info()["chains_per_thread"] == 4
    cnt, par, mean, cm, s2, ss = s.counters(), s.fetch("par"), s.fetch("mean"), s.fetch("cmat"), s.fetch("sigma2"), s.fetch("ss")
    assert (cnt["status"] == 0).all() and (cnt["simuind"] == steps + 1).all()
    oblob = O.blob_expreg(x, y)
    for c in (0, 77777, 148 * 512 * 4 - 1, N - 1):  # first tile, the bulk, the last 4-chain tile, the last small tile
        ch = oracle_chain(nml, O.MODEL_EXPREG, oblob, par0[c], cmat0, [0.5], [10000], 2024, c)
        ch.run()
        r = ch.results()
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k)
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-10)
        np.testing.assert_allclose(ss[c], r["sschain"][-1, 0], rtol=1e-10)
        np.testing.assert_allclose(s2[c], r["sigma2"], rtol=1e-10)
        np.testing.assert_allclose(mean[c], r["mean"], rtol=1e-10)
        iu = np.triu_indices(2)
        np.testing.assert_allclose(cm[c][iu], r["cmat"][iu], rtol=1e-7)
        assert r["drtries"] > 100 and r["chainind"] > 30
    s.close()


# ------------------------------------------------------------------------------------------------ C5 dimension
def test_scam_200_parameter_hierarchical_model():
    G, J, N = 198, 10, 4
    rng = np.random.default_rng(5)
    y = rng.normal(size=(G, 1)) + rng.normal(size=(G, J))
    d = G + 2
    blob = mb.models.blob_hier(y)
    # a covariance that stays full rank with well separated eigenvalues (eigenvectors of a degenerate
    # covariance are not unique: parity there is distributional)
    cmat0 = np.diag(0.01 * (1.0 + np.arange(d) / d))
    nml = dict(method="scam", nsimu=7, adaptint=3, initcmatn=400, updatesigma=0)
    par0 = 0.1 + 0.01 * rng.normal(size=(N, d))
    s = mb.Sampler(mb.default_config(nchains=N, seed=9, model="hier", **nml))
    s.set_data(blob)
    s.set_initial(par0, cmat0, [1.0], [1])
    s.run(6)
    cnt, par, q, ss = s.counters(), s.fetch("par"), s.fetch("qcovstd"), s.fetch("ss")
    assert (cnt["status"] == 0).all()
    for c in range(N):
        ch = oracle_chain(nml, O.MODEL_HIER, O.blob_hier(y), par0[c], cmat0, [1.0], [1], 9, c)
        ch.run()
        r = ch.results()
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k, cnt[k][c], r[k])
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(ss[c], r["sschain"][-1, 0], rtol=1e-10)
        np.testing.assert_allclose(q[c], r["qcovstd"], rtol=1e-8)
        assert r["ndrawn"] > 6 * d  # two ticks happened and all 200 components were proposed every sweep
    s.close()


# ------------------------------------------------------------------------------------------------ C4 shape
def test_pooled_ram_banana_d50():
    d, N = 50, 64
    blob = mb.models.blob_banana(d, 0.03)
    nml = dict(method="ram", nsimu=41, adaptint=20, updatesigma=0, alphatarget=0.234, nuparam=0.7)
    par0 = 0.1 * np.random.default_rng(6).normal(size=(N, d))
    s = mb.Sampler(mb.default_config(nchains=N, seed=3, model="banana", pool_adapt=1, **nml))
    s.set_data(blob)
    s.set_initial(par0, np.eye(d), [1.0], [1])
    s.run(40)
    chains, ticks = O.run_pooled(O.make_cfg(**nml), O.MODEL_BANANA, blob, par0, np.eye(d), [1.0], [1], seed=3)
    assert len(ticks) == 2
    cnt, par, R = s.counters(), s.fetch("par"), s.fetch("R")
    iu = np.triu_indices(d)
    bad = 0
    for k, ch in enumerate(chains):
        r = ch.results()
        same = all(cnt[key][k] == r[key] for key in CNT)
        bad += not same
        if same:
            np.testing.assert_allclose(par[k], r["par"], rtol=1e-7, atol=1e-9)
            np.testing.assert_allclose(R[k][iu], r["R"][iu], rtol=1e-6, atol=1e-9)
    assert bad <= 2  # dchdd amplifies rounding as ||a|| -> 1: a flip desynchronises a chain for good
    W, _, S = s.pool_fetch()
    assert W == N
    np.testing.assert_allclose(S, ticks[-1][3], rtol=1e-6, atol=1e-9 * np.abs(ticks[-1][3]).max())
    s.close()


# ------------------------------------------------------------------------------------------------ streamed dumps
@pytest.mark.parametrize("kernel", ["k1", "k2"])
def test_every_streamed_snapshot_matches_the_oracle(kernel):
    N, stride, nsnap = 24, 25, 4
    if kernel == "k1":
        nml = dict(cases.NML_DRAM, nsimu=stride * nsnap + 1, adaptint=50)
        args = ("expreg", BLOB11, np.tile(cases.PAR0, (N, 1)), cases.CMAT0, cases.SIGMA2, cases.NOBS)
        oid = O.MODEL_EXPREG
    else:
        d = 6
        mu, lam = gauss_target(d)
        nml = dict(nsimu=stride * nsnap + 1, adaptint=50, drscale=2.0, initcmatn=1, updatesigma=1, N0=4.0, S02=1.0)
        args = ("gauss", mb.models.blob_gauss(mu, lam), 0.1 * np.random.default_rng(8).normal(size=(N, d)), 0.05 * np.eye(d),
                [1.0], [3])
        oid = O.MODEL_GAUSS
    s = mb.Sampler(mb.default_config(nchains=N, seed=21, model=args[0], dump_stride=stride, **nml))
    s.set_data(args[1])
    s.set_initial(*args[2:])
    snaps = []
    for _ in range(nsnap):  # pop between runs: the ring holds 4 snapshots
        s.run(stride)
        got = s.dump_pop_ex()
        assert got is not None
        snaps.append(got)
    assert s.dump_pop_ex() is None
    assert [g[0] for g in snaps] == [1 + stride * (k + 1) for k in range(nsnap)]
    for c in range(N):
        ch = oracle_chain(nml, oid, args[1], args[2][c], args[3], args[4], args[5], 21, c)
        for step, par, ss, s2 in snaps:
            ch.advance(step)
            r = ch.results()
            assert r["simuind"] == step
            np.testing.assert_allclose(par[c], r["par"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(ss[c], r["sschain"][-1, :-1], rtol=1e-9)
            np.testing.assert_allclose(s2[c], r["sigma2"], rtol=1e-9)
    assert np.array_equal(snaps[-1][1], s.fetch("par"))
    s.close()


# ------------------------------------------------------------------------------------------------ restart files
def test_restart_files_and_continuation_run(tmp_path):
    leg1, leg2 = str(tmp_path / "leg1"), str(tmp_path / "leg2")
    os.makedirs(leg1)
    nml = NML_SHIPPED.replace("method = 'dram'", "method = 'dram'\n nmlffile = 'final.nml'\n covnfile = 'mcmccovn.dat'") \
                     .replace("drscale     = 0", "drscale     = 2.0").replace("burnintime  = 1000", "burnintime  = 0") \
                     .replace("doburnin    = 1", "doburnin    = 0").replace("nsimu       = 1000", "nsimu       = 300") \
                     .replace("adaptint    = 200", "adaptint    = 50\n initcmatn = 1")
    write_testcase(leg1, nml, "&mcmcb nchains = 2, seed = 17, store_chains = 1 /\n")
    # covnfile is read by initialize when it exists (initialize.F90:93-103): absent in leg 1
    r = subprocess.run([os.path.join(HOST, "mcmcb_main"), leg1], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    kw1 = dict(nsimu=300, doadapt=1, adaptint=50, initcmatn=1, burnintime=0, doburnin=0, drscale=2.0, updatesigma=1, N0=1.0,
               S02=0.0)
    ch = oracle_chain(kw1, O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), cases.PAR0, cases.CMAT0, cases.SIGMA2,
                      cases.NOBS, 17, 0)
    ch.run()
    ref = ch.results()
    # step 300 is an adaptation tick: the files hold chaincmat / chainmean / chainwsum as MCMC_writechains saves them
    covf = np.loadtxt(os.path.join(leg1, "mcmccovf.dat"))
    np.testing.assert_allclose(covf[np.triu_indices(2)], ref["cmat"][np.triu_indices(2)], rtol=1e-8)
    np.testing.assert_allclose(covf, covf.T, rtol=0, atol=0)
    np.testing.assert_allclose(np.loadtxt(os.path.join(leg1, "mcmcmean.dat")), ref["mean"], rtol=1e-10)
    assert float(np.loadtxt(os.path.join(leg1, "mcmccovn.dat"))) == float(int(ref["wsum"]))
    np.testing.assert_allclose(np.loadtxt(os.path.join(leg1, "mcmcparf.dat")), ref["par"], rtol=1e-10)
    text = open(os.path.join(leg1, "final.nml")).read()
    # MCMC_aux.F90:48-52,79-83: initcmatn = int(chainwsum) + simuind, burnintime = 0
    assert "initcmatn = %d," % (int(ref["wsum"]) + 300) in text and "burnintime = 0," in text

    # ---- leg 2: the user copies the *f.dat files over the inputs and runs with the written namelist
    shutil.copytree(leg1, leg2)
    for src, dst in (("mcmccovf.dat", "mcmccov.dat"), ("mcmcparf.dat", "mcmcpar.dat"), ("mcmcsigma2f.dat", "mcmcsigma2.dat"),
                     ("final.nml", "mcmcinit.nml")):
        shutil.copy(os.path.join(leg2, src), os.path.join(leg2, dst))
    n0 = int(np.loadtxt(os.path.join(leg2, "mcmccovn.dat")))  # initialize reads initcmatn from covnfile when it exists
    r = subprocess.run([os.path.join(HOST, "mcmcb_main"), leg2], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    par0 = np.loadtxt(os.path.join(leg2, "mcmcpar.dat"))
    cmat0 = np.loadtxt(os.path.join(leg2, "mcmccov.dat"))
    s2n = np.loadtxt(os.path.join(leg2, "mcmcsigma2.dat"))
    kw2 = dict(kw1, initcmatn=n0, S02=0.5)  # the written namelist carries S02 = sigma2(1) of leg 1 (MCMC_init.F90:114-116)
    ch2 = oracle_chain(kw2, O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), par0, cmat0, [s2n[0]], [int(s2n[1])], 17, 0)
    ch2.run()
    ref2 = ch2.results()
    chain = np.loadtxt(os.path.join(leg2, "chain.dat"), ndmin=2)
    assert chain.shape == ref2["chain"].shape and np.array_equal(chain[:, -1], ref2["chain"][:, -1])
    np.testing.assert_allclose(chain[:, :-1], ref2["chain"][:, :-1], rtol=1e-9)
    np.testing.assert_allclose(np.loadtxt(os.path.join(leg2, "mcmcmean.dat")), ref2["mean"], rtol=1e-9)
    assert float(np.loadtxt(os.path.join(leg2, "mcmccovn.dat"))) == float(int(ref2["wsum"]))


# ------------------------------------------------------------------------------------------------ long burn-in
def test_tick_kernels_with_a_row_buffer_beyond_48k():
    # rowcap = burnintime + 2 adaptint + ... rows: 16 bytes of per-row coefficients used to live in shared memory
    d, N = 6, 8
    mu, lam = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    nml = dict(nsimu=5301, adaptint=100, burnintime=5000, doburnin=1, scalelimit=0.05, initcmatn=1, updatesigma=0, drscale=0.0)
    par0 = 0.1 * np.random.default_rng(3).normal(size=(N, d))
    s = mb.Sampler(mb.default_config(nchains=N, seed=4, model="gauss", **nml))
    s.set_data(blob)
    s.set_initial(par0, 0.05 * np.eye(d), [1.0], [1])
    s.run(5300)
    cnt, par = s.counters(), s.fetch("par")
    assert (cnt["status"] == 0).all()
    for c in (0, N - 1):
        ch = oracle_chain(nml, O.MODEL_GAUSS, blob, par0[c], 0.05 * np.eye(d), [1.0], [1], 4, c)
        ch.run()
        r = ch.results()
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k)
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8, atol=1e-10)
    s.close()


# ------------------------------------------------------------------------------------------------ one handle, n GPUs
def test_group_handle_of_one_device_equals_plain_handle():
    nml = dict(cases.NML_DRAM, nsimu=201)
    out = []
    for g in (0, 1):
        s = mb.Sampler(mb.default_config(nchains=100, seed=5, ngpus=g, **nml))
        s.set_data(BLOB11)
        s.set_initial(cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS)
        s.run(200)
        out.append((s.fetch("par"), s.fetch("counters")))
        assert s.ngpus == 1
        s.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("pool", [0, 1])
def test_one_handle_drives_two_gpus(pool):
    if ndev() < 2:
        pytest.skip("needs 2 GPUs on the box (run under `gpurun --gpus 2`)")
    N = 1001  # odd: the shards differ in size
    nml = dict(cases.NML_DRAM, nsimu=301, adaptint=50)
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(1).normal(size=(N, 2)))

    def run(ngpus):
        s = mb.Sampler(mb.default_config(nchains=N, seed=3, ngpus=ngpus, pool_adapt=pool, store_chains=N, dump_stride=100, **nml))
        s.set_data(BLOB11)
        s.set_initial(par0, cases.CMAT0, cases.SIGMA2, cases.NOBS)
        s.run(300)
        snaps = []
        while True:
            g = s.dump_pop_ex()
            if g is None:
                break
            snaps.append(g)
        res = dict(par=s.fetch("par"), cnt=s.fetch("counters"), cm=s.fetch("cmat"), R=s.fetch("R"), snaps=snaps,
                   chain=s.fetch_chain(N - 1)["chain"], ngpus=s.ngpus, nccl=s.nccl_calls)
        if pool:
            res["pool"] = s.pool_fetch()
        s.close()
        return res

    one, two = run(1), run(2)
    assert one["ngpus"] == 1 and two["ngpus"] == 2
    assert [g[0] for g in two["snaps"]] == [101, 201, 301]
    if not pool:  # no collective: results do not depend on the number of devices, bit for bit
        assert two["nccl"] == 0
        for k in ("par", "cnt", "cm", "R", "chain"):
            assert np.array_equal(one[k], two[k]), k
        for a, b in zip(one["snaps"], two["snaps"]):
            assert all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))
    else:  # pooled ticks: 6 ticks x 2 allreduces over NCCL; block-partial sums are formed per device -> rounding level
        assert two["nccl"] == 12
        assert np.array_equal(one["cnt"][:, :6], two["cnt"][:, :6])
        np.testing.assert_allclose(one["par"], two["par"], rtol=1e-9)
        np.testing.assert_allclose(one["pool"][2], two["pool"][2], rtol=1e-10)
        assert one["pool"][0] == two["pool"][0]
        assert np.array_equal(two["R"], np.broadcast_to(two["R"][0], two["R"].shape))


# ------------------------------------------------------------------------------------------------ K1 beyond ExpReg
@pytest.mark.parametrize("d,method", [(5, "dram"), (5, "ram"), (3, "dram"), (8, "dram"), (7, "dram"), (5, "scam")])
def test_small_npar_gaussian_runs_on_the_register_kernel(d, method):
    """testcases/mcmcrun4.F90 is a 5-parameter Gaussian target: "gauss" has compile-time-npar registrations for the
    register kernel (3, 4, 5, 6, 8); other npar, and samplers that need an SVD factor, take the warp-per-chain kernels."""
    mu, lam = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    N = 40
    if method == "ram":
        nml = dict(method="ram", nsimu=201, updatesigma=1, N0=3.0, S02=1.0)
    elif method == "scam":
        nml = dict(method="scam", nsimu=201, adaptint=50, initcmatn=2 * d, updatesigma=0)
    else:
        nml = dict(nsimu=201, adaptint=50, drscale=2.0, initcmatn=1, updatesigma=1, N0=3.0, S02=1.0)
    par0 = 0.1 * np.random.default_rng(d).normal(size=(N, d))
    cmat0 = np.diag(0.05 * (1.0 + np.arange(d)))
    s = mb.Sampler(mb.default_config(nchains=N, seed=31, model="gauss", **nml))
    s.set_data(blob)
    s.set_initial(par0, cmat0, [1.0], [4])
    s.run(200)
    assert s.info()["kernel"] == (1 if d in (3, 4, 5, 6, 8) and method != "scam" else 2)
    cnt, par, s2 = s.counters(), s.fetch("par"), s.fetch("sigma2")
    assert (cnt["status"] == 0).all()
    for c in (0, N - 1):
        ch = oracle_chain(nml, O.MODEL_GAUSS, blob, par0[c], cmat0, [1.0], [4], 31, c)
        ch.run()
        r = ch.results()
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k)
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8 if method != "dram" else 1e-10, atol=1e-12)
        np.testing.assert_allclose(s2[c], r["sigma2"], rtol=1e-9)
    st = s.fetch_stats(N - 1)  # the typed single-chain fetch agrees with the string-keyed population fetch
    assert np.array_equal(st["mean"], s.fetch("mean")[N - 1]) and st["wsum"] == s.fetch("wsum")[N - 1, 0]
    assert np.array_equal(st["cmat"], s.fetch("cmat")[N - 1]) and st["counters"]["stayed"] == cnt["stayed"][N - 1]
    assert np.array_equal(np.triu(st["R"]), np.triu(s.fetch("R")[N - 1]))
    s.close()


# ------------------------------------------------------------------------------------------------ thread-per-chain RAM
@pytest.mark.parametrize("case", ["banana50", "gauss6_stored", "bounds"])
def test_thread_per_chain_ram_kernel_matches_the_oracle(case, monkeypatch):
    """k4_ram_step_kernel (one thread per chain, packed SoA factors) is what large RAM populations run on (BASELINE C4);
    forced here on a small population and compared with the oracle and with the warp-per-chain kernel."""
    if case == "banana50":
        d, N, steps = 50, 96, 80
        model, oid, blob = "banana", O.MODEL_BANANA, mb.models.blob_banana(50, 0.03)
        nml = dict(method="ram", nsimu=steps + 1, updatesigma=0, alphatarget=0.234, nuparam=0.7)
        par0, cmat0, s2, nobs = 0.1 * np.random.default_rng(1).normal(size=(N, d)), np.eye(d), [1.0], [1]
    elif case == "gauss6_stored":
        d, N, steps = 6, 40, 200
        mu, lam = gauss_target(d)
        model, oid, blob = "gauss", O.MODEL_GAUSS, mb.models.blob_gauss(mu, lam)
        nml = dict(method="ram", nsimu=steps + 1, updatesigma=1, N0=3.0, S02=1.0, burnintime=20, doburnin=1)
        par0, cmat0, s2, nobs = 0.1 * np.random.default_rng(2).normal(size=(N, d)), 0.3 * np.eye(d), [1.0], [4]
    else:  # out-of-bounds proposals: the stale alpha12 drives the adaptation (Q11)
        d, N, steps = 2, 40, 150
        model, oid, blob = "expreg", O.MODEL_EXPREG, BLOB11
        nml = dict(method="ram", nsimu=steps + 1, updatesigma=1, N0=1.0, S02=0.0)
        par0, cmat0, s2, nobs = np.tile([10.0, 0.02], (N, 1)), np.diag([0.2, 0.004]), cases.SIGMA2, cases.NOBS
    out = {}
    for k4 in ("1", "0"):
        monkeypatch.setenv("MCMCB_K4", k4)
        s = mb.Sampler(mb.default_config(nchains=N, seed=13, model=model, kernel=2, store_chains=2, **nml))
        s.set_data(blob)
        s.set_initial(par0, cmat0, s2, nobs)
        s.run(steps // 2)
        s.run(steps - steps // 2)  # two launches: the factors survive the pack / unpack round trip
        assert s.info()["lanes_per_chain"] == (1 if k4 == "1" else 32)
        out[k4] = dict(cnt=s.counters(), par=s.fetch("par"), R=s.fetch("R"), s2=s.fetch("sigma2"), chain=s.fetch_chain(1))
        s.close()
    a = out["1"]
    iu = np.triu_indices(d)
    for c in (0, 1, N - 1):
        ch = oracle_chain(nml, oid, blob, par0[c], cmat0, s2, nobs, 13, c)
        ch.run()
        r = ch.results()
        for k in CNT + ("status",):  # a failed downdate (status 2, matutils.F90:716-722) is flagged by both, same chains
            assert a["cnt"][k][c] == r[k], (c, k)
        np.testing.assert_allclose(a["par"][c], r["par"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(a["R"][c][iu], r["R"][iu], rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(a["s2"][c], r["sigma2"], rtol=1e-9)
        if c == 1:
            assert np.array_equal(a["chain"]["chain"][:, -1], r["chain"][:, -1])
            np.testing.assert_allclose(a["chain"]["chain"][:, :-1], r["chain"][:, :-1], rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(a["chain"]["s2chain"], r["s2chain"], rtol=1e-9)
    if case == "bounds":
        assert a["cnt"]["bndstayed"].sum() > 0
    else:
        assert (a["cnt"]["status"] == 0).all()
    # one launch instead of two: bit-identical (the look-ahead proposal of the fused sweeps does not leak across launches)
    monkeypatch.setenv("MCMCB_K4", "1")
    s = mb.Sampler(mb.default_config(nchains=N, seed=13, model=model, kernel=2, store_chains=2, **nml))
    s.set_data(blob)
    s.set_initial(par0, cmat0, s2, nobs)
    s.run(steps)
    assert np.array_equal(s.fetch("par"), a["par"]) and np.array_equal(s.fetch("R"), a["R"])
    assert np.array_equal(s.fetch("counters"), np.column_stack([a["cnt"][k] for k in mb.binding.COUNTER_NAMES]))
    s.close()
    # the two kernels walk the same chains (values differ at rounding level: lane-strided sums in the warp kernel)
    b = out["0"]
    same = sum(all(a["cnt"][k][c] == b["cnt"][k][c] for k in CNT) for c in range(N))
    assert same >= N - 2


# ------------------------------------------------------------------------------------------------ thread-per-chain SCAM
@pytest.mark.parametrize("lanes", [4, 2, 1, 0])
@pytest.mark.parametrize("G,J,N,steps", [(6, 4, 64, 59), (198, 10, 8, 3), (13, 3, 150, 21)])
def test_thread_per_chain_scam_kernel_matches_the_oracle(G, J, N, steps, lanes, monkeypatch):
    """k5_scam_step_kernel (one thread per chain, shared rotation, proposal never materialised: HierN::ssfunction_axpy)
    is what large pooled SCAM populations run on (BASELINE C5); forced here on a small population.  Up to the first
    pooled tick every chain is the reference's own SCAM chain: exact counters against the oracle."""
    rng = np.random.default_rng(G)
    y = rng.normal(size=(G, 1)) + rng.normal(size=(G, J))
    blob = mb.models.blob_hier(y)
    d = G + 2
    nml = dict(method="scam", nsimu=steps + 1, adaptint=steps + 1, initcmatn=1, updatesigma=1, N0=2.0, S02=1.0)
    par0 = 0.1 + 0.05 * rng.normal(size=(N, d))
    cmat0 = np.diag(0.02 * (1.0 + np.arange(d) / d))
    monkeypatch.setenv("MCMCB_K5", "1")
    # theta in shared memory, 4 / 2 / 1 lanes per chain (k5s_scam.cuh); 0: theta in local memory (k5_scam.cuh)
    monkeypatch.setenv("MCMCB_K5S", "1" if lanes else "0")
    monkeypatch.setenv("MCMCB_K5S_LANES", str(max(lanes, 1)))
    s = mb.Sampler(mb.default_config(nchains=N, seed=23, model="hier", pool_adapt=1, store_chains=2, **nml))
    s.set_data(blob)
    s.set_initial(par0, cmat0, [1.0], [G * J])
    s.run(steps // 2)
    s.run(steps - steps // 2)
    assert s.info()["lanes_per_chain"] == max(lanes, 1)
    assert (s.info()["smem_bytes"] > 8 * d * 64) == (lanes > 0)
    cnt, par, ss, s2 = s.counters(), s.fetch("par"), s.fetch("ss"), s.fetch("sigma2")
    assert (cnt["status"] == 0).all()
    for c in (0, 1, N - 1):
        ch = oracle_chain(nml, O.MODEL_HIER, blob, par0[c], cmat0, [1.0], [G * J], 23, c)
        ch.run()
        r = ch.results()
        for k in CNT:
            assert cnt[k][c] == r[k], (c, k, cnt[k][c], r[k])
        np.testing.assert_allclose(par[c], r["par"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(ss[c], r["sschain"][-1, 0], rtol=1e-10)
        np.testing.assert_allclose(s2[c], r["sigma2"], rtol=1e-9)
        if c == 1:
            g = s.fetch_chain(1)
            assert np.array_equal(g["chain"][:, -1], r["chain"][:, -1])
            np.testing.assert_allclose(g["chain"][:, :-1], r["chain"][:, :-1], rtol=1e-8, atol=1e-10)
    s.close()


def test_thread_per_chain_scam_kernel_across_pooled_ticks(monkeypatch):
    """Across pooled ticks both step kernels take their rotation from the same device SVD of the pooled covariance: the
    thread-per-chain kernel walks the chains of the warp-per-chain kernel (sums differ at rounding level)."""
    G, J, N = 6, 4, 96
    rng = np.random.default_rng(7)
    y = rng.normal(size=(G, 1)) + rng.normal(size=(G, J))
    blob = mb.models.blob_hier(y)
    d = G + 2
    nml = dict(method="scam", nsimu=151, adaptint=50, initcmatn=1, updatesigma=0)
    par0 = 0.1 * rng.normal(size=(N, d))
    out = {}
    for k5 in ("1", "0"):
        monkeypatch.setenv("MCMCB_K5", k5)
        s = mb.Sampler(mb.default_config(nchains=N, seed=3, model="hier", pool_adapt=1, **nml))
        s.set_data(blob)
        s.set_initial(par0, 0.1 * np.eye(d), [1.0], [1])
        s.run(150)
        assert s.info()["lanes_per_chain"] == (4 if k5 == "1" else 32)   # 4 lanes per chain: k5s_scam.cuh
        out[k5] = dict(cnt=s.counters(), par=s.fetch("par"), pool=s.pool_fetch(), q=s.fetch("qcovstd"))
        assert (out[k5]["cnt"]["status"] == 0).all()
        s.close()
    a, b = out["1"], out["0"]
    same = sum(all(a["cnt"][k][c] == b["cnt"][k][c] for k in CNT) for c in range(N))
    assert same >= N - 4, same
    assert a["pool"][0] == b["pool"][0]
    np.testing.assert_allclose(a["pool"][1], b["pool"][1], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(a["q"][0], b["q"][0], rtol=1e-5)


# ------------------------------------------------------------------------------------------------ resident tick
def _tick_run(resident, monkeypatch, model, blob, d, N, nml, par0, cmat0, pool=0, kernel=2, runs=()):
    monkeypatch.setenv("MCMCB_TICK_RESIDENT", "1" if resident else "0")
    s = mb.Sampler(mb.default_config(nchains=N, seed=11, model=model, kernel=kernel, pool_adapt=pool, **nml))
    s.set_data(blob)
    s.set_initial(par0, cmat0, [1.0], [1])
    for n in runs:
        s.run(n)
    out = dict(par=s.fetch("par"), cnt=s.counters(), launches=s.launches,
               stats=[s.fetch_stats(c) for c in (0, N // 2, N - 1)])
    if pool:
        out["pool"] = s.pool_fetch()
    s.close()
    return out


@pytest.mark.parametrize("d,method,initcmatn,pool", [(100, "dram", 1, 0), (18, "dram", 0, 0), (33, "dram", 1, 0),
                                                     (64, "scam", 1, 0), (24, "dram", 1, 1)])
def test_resident_tick_is_bit_identical_to_the_streamed_tick(d, method, initcmatn, pool, monkeypatch):
    """k2_absorb_resident_kernel (all logged rows in shared memory, 4 x 4 register tiles) against cta_absorb_rows (8-row
    window, runs of 8 entries): the covariance recursion of matutils.F90:283-310 applied to every entry in the same
    order, so means, covariances, factors, chains and counters must agree to the last bit -- ragged npar (18, 33),
    the wsum = 0 reset row (initcmatn = 0), the SVD factor modes and the pooled tick included."""
    N = 40
    mu, lam = gauss_target(d)
    blob = mb.models.blob_gauss(mu, lam)
    a = 2 * d if not pool else 30
    nml = dict(nsimu=3 * a + 10, adaptint=a, initcmatn=initcmatn, updatesigma=0, method=method,
               drscale=2.0 if method == "dram" else 0.0)
    par0 = 0.1 * np.random.default_rng(d).normal(size=(N, d))
    runs = (a + 3, a, a + 2)   # three ticks, launches split off the tick boundaries
    got = [_tick_run(r, monkeypatch, "gauss", blob, d, N, nml, par0, 0.02 * np.eye(d), pool=pool, runs=runs) for r in (0, 1)]
    if not pool:                                            # the resident kernel did run: one more launch per private tick
        assert got[1]["launches"] == got[0]["launches"] + 3
    assert np.array_equal(got[0]["par"], got[1]["par"])
    for k in CNT + ("status",):
        assert np.array_equal(got[0]["cnt"][k], got[1]["cnt"][k]), k
    for x, y in zip(got[0]["stats"], got[1]["stats"]):
        assert x["wsum"] == y["wsum"] and x["wsum"] > 0
        for k in ("mean", "cmat", "R"):
            assert np.array_equal(x[k], y[k]), k
        assert np.array_equal(x["cmat"], x["cmat"].T)
    if pool:
        for x, y in zip(got[0]["pool"], got[1]["pool"]):
            assert np.array_equal(np.asarray(x), np.asarray(y))


# ------------------------------------------------------------------------------------------------ theta in shared memory
@pytest.mark.parametrize("lanes,cpt", [(1, 1), (2, 1), (4, 1), (4, 2), (8, 2), (8, 4)])
def test_scam_kernel_with_theta_in_shared_memory(lanes, cpt, monkeypatch):
    """k5s_scam_step_kernel against k5_scam_step_kernel: the same draws and, element by element, the same fma sequence
    (theta + delta U(:,j) composed on the fly, rewritten on acceptance).  With one lane per chain the model's sum runs in
    the same order: chains, sums of squares, counters, logged rows (through the pooled covariance of three ticks) and
    the stored chain agree to the last bit.  With 2 or 4 lanes the sum of squares is the sum of the lanes' partial sums
    (rounding-level differences, as in the warp-per-chain kernel): same walks, values to 1e-9.  150 chains: the last
    CTA has shadow slots; 198 groups: ragged last blocks of the model's view.  cpt = chains per thread (the model's
    batched view: one sweep over the data for all of them) does not change a single bit."""
    G, J, N = 198, 3, 150
    rng = np.random.default_rng(31)
    y = rng.normal(size=(G, 1)) + rng.normal(size=(G, J))
    blob = mb.models.blob_hier(y)
    d = G + 2
    nml = dict(method="scam", nsimu=14, adaptint=4, initcmatn=1, updatesigma=1, N0=2.0, S02=1.0)
    par0 = 0.1 * rng.normal(size=(N, d))
    monkeypatch.setenv("MCMCB_K5", "1")
    monkeypatch.setenv("MCMCB_K5S_LANES", str(lanes))
    out = {}
    for k5s in ("1", "0", "1x"):
        if k5s == "1x" and cpt == 1:
            continue
        monkeypatch.setenv("MCMCB_K5S", k5s[0])
        monkeypatch.setenv("MCMCB_K5S_CPT", "1" if k5s == "1x" else str(cpt))
        s = mb.Sampler(mb.default_config(nchains=N, seed=5, model="hier", pool_adapt=1, store_chains=3, **nml))
        s.set_data(blob)
        s.set_initial(par0, 0.05 * np.eye(d), [1.0], [G * J])
        s.run(6)
        s.run(7)
        out[k5s] = dict(cnt=s.counters(), par=s.fetch("par"), ss=s.fetch("ss"), s2=s.fetch("sigma2"), pool=s.pool_fetch(),
                        chain=s.fetch_chain(2), smem=s.info()["smem_bytes"], lanes=s.info()["lanes_per_chain"])
        assert (out[k5s]["cnt"]["status"] == 0).all()
        s.close()
    a, b = out["1"], out["0"]
    assert a["smem"] > 8 * d * 64 > b["smem"] and a["lanes"] == lanes and b["lanes"] == 1
    if "1x" in out and lanes == 4:   # (lanes 8 has no one-chain-per-thread instantiation: MCMCB_K5S_CPT=1 is ignored there)
        x = out["1x"]
        for k in CNT:
            assert np.array_equal(a["cnt"][k], x["cnt"][k]), k
        for k in ("par", "ss", "s2"):
            assert np.array_equal(a[k], x[k]), k
    if lanes == 1:
        for k in CNT:
            assert np.array_equal(a["cnt"][k], b["cnt"][k]), k
        for k in ("par", "ss", "s2"):
            assert np.array_equal(a[k], b[k]), k
        for x, y_ in zip(a["pool"], b["pool"]):
            assert np.array_equal(np.asarray(x), np.asarray(y_))
        for k in ("chain", "sschain", "s2chain"):
            assert np.array_equal(a["chain"][k], b["chain"][k]), k
    else:
        same = [c for c in range(N) if all(a["cnt"][k][c] == b["cnt"][k][c] for k in CNT)]
        assert len(same) >= N - 3, len(same)
        np.testing.assert_allclose(a["par"][same], b["par"][same], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(a["ss"][same], b["ss"][same], rtol=1e-11)
        assert a["pool"][0] == b["pool"][0]
        if len(same) == N:
            np.testing.assert_allclose(a["pool"][2], b["pool"][2], rtol=1e-7, atol=1e-10)

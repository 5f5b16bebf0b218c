"""Pins of the C oracle (oracle/mcmc_oracle.c) to the reference, without a Fortran compiler.

Neither this container nor the GPU box has one (profiles/r02_probe_fortran.txt), so:

1. `oracle/restate_np.py` -- a second, independent reading of the Fortran (array-at-a-time numpy, REAL BLAS/LAPACK
   through scipy instead of the C oracle's restated netlib loops) -- is run on the same injected uniform streams as
   the C oracle for every sampler the reference has (DRAM/AM as shipped, DRAM with DR, RAM, SCAM, ER, AP window,
   greedy burn-in; the shipped testcase and larger Gaussian / hierarchical targets).  Chain indices, repeat counts,
   every counter and the number of uniforms consumed must be EQUAL; chain values, ss, sigma2, covariance and factor
   must agree to 1e-9 relative (the two use different BLAS operation orders and FMA contraction).
2. The one binary fixture the reference holds, testcases/data.mat (copied to tests/golden/ref_testcases_data.mat),
   byte-pins the MAT-v4 writer: writing testcases/data.dat's values under the name "data" must reproduce it.
3. If a Fortran compiler exists, oracle/_ref/Makefile builds the reference itself with its random_number calls
   reading the injected stream, and the oracle is compared with its chain files (skipped, loudly, otherwise).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import restate_np as R2
from tests import cases
from tests.golden import make_golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def np_model(name):
    if name.startswith("gauss_"):
        mu, lam = G.gauss_target(G.GAUSS_D)
        return R2.Gauss(mu, lam)
    if name == "scam_hier":
        return R2.Hier(G.HIER_Y)
    return R2.ExpReg(cases.DATA_X, cases.DATA_Y)


def run_both(nml, model_id, blob, model, par0, cmat0, sigma2, nobs, u):
    ch = O.Chain(O.make_cfg(**nml), model_id, blob, par0, cmat0, sigma2, nobs)
    ch.inject(u)
    ch.run()
    a = ch.results()
    b = R2.Run(dict(nml), model, par0, cmat0, sigma2, nobs, u).run()
    return a, b


def compare(a, b, rtol=1e-9, svd=False):
    cb = b.counters()
    for k in ("stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "erstayed", "ndrawn"):
        assert a[k] == cb[k], (k, a[k], cb[k])
    n = a["chainind"]
    d = b.npar
    # repeat counts (the chain "indices" of parity check 1) exactly
    assert np.array_equal(a["chain"][:, d], b.chain[:n, d])
    assert np.array_equal(a["sschain"][:, -1], b.sschain[:n, -1])
    scale = np.abs(b.chain[:n, :d]).max(axis=0)
    assert np.allclose(a["chain"][:, :d], b.chain[:n, :d], rtol=rtol, atol=rtol * scale.max())
    assert np.allclose(a["sschain"][:, :-1], b.sschain[:n, :-1], rtol=rtol)
    if b.updatesigma:
        assert np.allclose(a["s2chain"], b.s2chain, rtol=rtol)
    assert a["wsum"] == b.chainwsum
    assert np.allclose(a["mean"], b.chainmean, rtol=1e-8, atol=1e-8 * scale.max())
    cm_a = np.triu(a["cmat"]) + np.triu(a["cmat"], 1).T
    assert np.allclose(cm_a, b.chaincmat, rtol=1e-7, atol=1e-7 * np.abs(b.chaincmat).max())
    if svd:
        # factor of a symmetric matrix with well separated eigenvalues: same up to rounding after the sign convention
        assert np.allclose(a["R"], b.R, rtol=1e-5, atol=1e-6 * np.abs(b.R).max())
        if b.doscam:
            assert np.allclose(a["qcovstd"], b.qcovstd, rtol=1e-7)
    else:
        assert np.allclose(np.triu(a["R"]), np.triu(b.R), rtol=1e-7, atol=1e-9 * np.abs(b.R).max())
        if b.dodr:
            assert np.allclose(np.triu(a["iC"]), np.triu(b.iC), rtol=1e-6)
            assert np.allclose(np.triu(a["R2"]), np.triu(b.R2), rtol=1e-7, atol=1e-9 * np.abs(b.R).max())


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_equals_independent_restatement(name):
    """Every golden case (tests/golden/make_golden.py): the reference's shipped testcase under all samplers and the
    6-dimensional Gaussian / hierarchical targets."""
    model_id, blob, par0, cmat0, sigma2, nobs = G.inputs(name)
    a, b = run_both(G.CASES[name], model_id, blob, np_model(name), par0, cmat0, sigma2, nobs, G.uniforms(name))
    compare(a, b, svd=name.startswith("scam"))
    assert a["chainind"] > 10  # the case really moves


def test_shipped_namelist_full_length():
    """testcases/mcmcinit.nml verbatim (nsimu = 1000, burn-in scaling ticks every 200 steps, sigma2 Gibbs)."""
    u = np.random.default_rng(101).random(60000)
    a, b = run_both(cases.NML_SHIPPED, O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y),
                    R2.ExpReg(cases.DATA_X, cases.DATA_Y), cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS, u)
    compare(a, b)


def test_dram_long_with_prior_and_bounds():
    """DRAM + AM over 20 adaptation ticks with the default Gaussian prior (priorfun.f90:97-100) and a start close
    to the bound so that out-of-bounds proposals occur (Q3 counters)."""
    nml = dict(cases.NML_DRAM, nsimu=2001)
    u = np.random.default_rng(102).random(200000)
    par0 = np.array([10.0, 0.02])
    cmat0 = np.diag([0.2, 0.004])
    prior = (np.array([10.0, 0.1]), np.array([2.0, 0.0]))
    ch = O.Chain(O.make_cfg(**nml), O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), par0, cmat0, cases.SIGMA2,
                 cases.NOBS, prior=prior)
    ch.inject(u)
    ch.run()
    a = ch.results()
    b = R2.Run(dict(nml), R2.ExpReg(cases.DATA_X, cases.DATA_Y), par0, cmat0, cases.SIGMA2, cases.NOBS, u, prior=prior).run()
    assert a["bndstayed"] > 0
    compare(a, b)


@pytest.mark.parametrize("d", [20, 50])
def test_ram_large_dim(d):
    """RAM (dchud / dchdd every step) on a d-dimensional correlated Gaussian: BASELINE C4's dimension."""
    mu, lam = G.gauss_target(d)
    nml = dict(method="ram", nsimu=201, updatesigma=0)
    u = np.random.default_rng(103 + d).random(40 * 201 * d)
    a, b = run_both(nml, O.MODEL_GAUSS, O.blob_gauss(mu, lam), R2.Gauss(mu, lam), np.zeros(d), 0.05 * np.eye(d), [1.0], [1], u)
    # dchdd amplifies last-bit differences as ||a|| -> 1: values to 1e-7, counts exactly
    compare(a, b, rtol=1e-7)


def test_dram_large_dim_two_ticks():
    """DRAM with DR at d = 30 over two AM ticks: dpotrf/dpotri through real LAPACK against the restated dpotf2/dtrti2/dlauu2."""
    d = 30
    mu, lam = G.gauss_target(d)
    nml = dict(nsimu=121, adaptint=50, drscale=2.0, initcmatn=40, updatesigma=0)
    u = np.random.default_rng(104).random(30 * 121 * d)
    a, b = run_both(nml, O.MODEL_GAUSS, O.blob_gauss(mu, lam), R2.Gauss(mu, lam), np.zeros(d), 0.05 * np.eye(d), [1.0], [1], u)
    compare(a, b, rtol=1e-8)


@pytest.mark.parametrize("nml", [
    dict(nsimu=301, adaptint=50, drscale=0.0, condmax=1e12, initcmatn=1, updatesigma=0),                 # AM, usesvd
    dict(nsimu=301, adaptint=40, adapthist=60, drscale=0.0, condmax=1e10, initcmatn=1, updatesigma=0),   # AP window -> SVD
    dict(method="scam", nsimu=201, adaptint=40, adapthist=60, initcmatn=1, updatesigma=0),               # SCAM + AP window
], ids=["am_usesvd", "ap_usesvd", "scam_ap"])
def test_svd_factor_modes_on_a_correlated_gaussian(nml):
    """covtor_svd / scam_svd (matutils.F90:378-453, 583-653) behind AM with condmax > 0 and behind SCAM, fed by the plain
    recursion and by the AP window (MCMC_adapt.F90:116-136): the C oracle's restated Jacobi against LAPACK's dgesvd."""
    d = 6
    mu, lam = G.gauss_target(d)
    u = np.random.default_rng(7).random(20 * 301 * d * 2)
    a, b = run_both(nml, O.MODEL_GAUSS, O.blob_gauss(mu, lam), R2.Gauss(mu, lam), 0.1 * np.ones(d), 0.05 * np.eye(d), [1.0], [1], u)
    assert b.usesvd and a["chainind"] > 100
    compare(a, b, svd=True)


def test_mat4_writer_reproduces_reference_fixture(tmp_path):
    """testcases/data.mat is the reference's own MAT-v4 image of testcases/data.dat (header 0,11,2,0,5,"data\\0",
    column-major float64; matfiles.F90:41-48,66-126): the writer must reproduce it byte for byte."""
    want = open(os.path.join(ROOT, "tests", "golden", "ref_testcases_data.mat"), "rb").read()
    assert len(want) == 201
    L = C.CDLL(os.path.join(ROOT, "host", "libmcmcbhost.so"))
    L.mcmcbh_write_mat4.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int]
    x = np.asfortranarray(np.column_stack([cases.DATA_X, cases.DATA_Y]))
    path = str(tmp_path / "data.mat")
    assert L.mcmcbh_write_mat4(path.encode(), b"data", x.ctypes.data_as(C.POINTER(C.c_double)), 11, 2, 11) == 0
    assert open(path, "rb").read() == want


def test_reference_build_when_a_fortran_compiler_exists(tmp_path):
    """oracle/_ref: the reference's own sources compiled with random_number redirected to the injected stream."""
    fc = next((f for f in ("gfortran", "gfortran-13", "gfortran-12", "flang", "ifx", "nvfortran") if shutil.which(f)), None)
    if fc is None:
        pytest.skip("NO FORTRAN COMPILER in this image (probed gfortran/flang/ifx/nvfortran here and on the GPU box, "
                    "profiles/r02_probe_fortran.txt): the reference itself cannot run; oracle pinned by restate_np only")
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference sources not present on this box")
    refdir = os.path.join(ROOT, "oracle", "_ref")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "FC=" + fc])
    for name in ("shipped", "dram", "ram", "er"):
        work = tmp_path / name
        work.mkdir()
        subprocess.check_call(["python", os.path.join(ROOT, "oracle", "ref_case.py"), name, str(work), os.path.join(refdir, "mcmcrun_ref")])
        chain = np.loadtxt(work / "chain.dat", ndmin=2)
        model_id, blob, par0, cmat0, sigma2, nobs = G.inputs(name)
        ch = O.Chain(O.make_cfg(**G.CASES[name]), model_id, blob, par0, cmat0, sigma2, nobs)
        ch.inject(G.uniforms(name))
        ch.run()
        a = ch.results()
        assert chain.shape[0] == a["chainind"]
        assert np.array_equal(chain[:, -1], a["chain"][:, -1])
        assert np.allclose(chain[:, :-1], a["chain"][:, :-1], rtol=1e-9)

"""Host side above the C ABI (host/): namelist &mcmc reader, `initialize` on the .dat files, the chain /
restart writers (ASCII and MAT-v4), and the driver executable.

not gpu: formats and parsing against the reference's shipped inputs (testcases/mcmcinit.nml etc., restated here
         because /root/reference does not travel) and against scipy.io (an independent MAT-v4 reader);
gpu:     `host/mcmcb_main` run in a copy of the testcase directory writes chain.dat / sschain.dat / s2chain.dat
         that equal the oracle's chain for the same Philox stream."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.io

import mcmcf90_b200 as mb
from mcmcf90_b200.binding import Config
from oracle import oracle as O
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")

# testcases/mcmcinit.nml as shipped
NML_SHIPPED = """!! 
!! Run time parameters for the mcmc run
!!
&mcmc
method = 'dram'
 nsimu       = 1000   ! length of the chain
 verbosity   = 1       ! how much to print
 doadapt     = 1       ! do we adapt
 adaptint    = 200     ! intervall for adaptation
 burnintime  = 1000    ! initial burn in time
 doburnin    = 1
 drscale     = 0       ! scaling factor for second stage DR
 printint    = 100     ! interval to print statistics
 updatesigma = 1       ! update error variance? (1=yes)
 N0          = 1       ! prior for error variance,
 S02         = 0       !   1/s^2 ~ Gamma(N0/2,2/N/S02)
 chainfile   = 'chain.dat'    ! file to save the chain
 ssfile      = 'sschain.dat'  ! save ssfunction values here
 s2file      = 's2chain.dat'  ! file to save sigma2 chain
! priorsfile = 'priors.dat'
/
"""


class Files(C.Structure):
    _fields_ = [(n, C.c_char * 256) for n in ("chainfile", "s2file", "ssfile", "priorsfile", "cov0file", "covffile",
                                              "covnfile", "meanfile", "nmlffile", "parfile", "parffile", "sigma2file",
                                              "sigma2ffile", "datafile")] + \
               [(n, C.c_int) for n in ("verbosity", "printint", "dumpint", "usrfunlen", "filepars", "svddim", "sstype")] + \
               [(n, C.c_double) for n in ("condmaxini", "sstrans")]


def hostlib():
    L = C.CDLL(os.path.join(HOST, "libmcmcbhost.so"))
    L.mcmcbh_last_error.restype = C.c_char_p
    dp = C.POINTER(C.c_double)
    L.mcmcbh_load_dat.argtypes = [C.c_char_p, C.POINTER(dp), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mcmcbh_write_dat.argtypes = [C.c_char_p, dp, C.c_int, C.c_int, C.c_int]
    L.mcmcbh_write_mat4.argtypes = [C.c_char_p, C.c_char_p, dp, C.c_int, C.c_int, C.c_int]
    L.mcmcbh_read_namelist.argtypes = [C.c_char_p, C.POINTER(Config), C.POINTER(Files)]
    L.mcmcbh_free.argtypes = [C.c_void_p]
    L.mcmcbh_addto_mat4.argtypes = [C.c_char_p, dp, C.c_int, C.c_int, C.c_int]
    L.mcmcbh_write_namelist.argtypes = [C.c_char_p, C.POINTER(Config), C.POINTER(Files)]
    return L


def write_testcase(d, nml=NML_SHIPPED, extra=""):
    open(os.path.join(d, "mcmcinit.nml"), "w").write(nml + extra)
    open(os.path.join(d, "mcmcpar.dat"), "w").write("10 0.1 \n")
    open(os.path.join(d, "mcmccov.dat"), "w").write("0.2 0 \n0 0.001 \n")
    open(os.path.join(d, "mcmcsigma2.dat"), "w").write("0.5\n11")
    with open(os.path.join(d, "data.dat"), "w") as f:
        f.write("% example data set\n")
        for x, y in zip(cases.DATA_X, cases.DATA_Y):
            f.write("   %d   %.2f\n" % (x, y))


def test_namelist_shipped_and_defaults(tmp_path):
    L = hostlib()
    p = tmp_path / "mcmcinit.nml"
    p.write_text(NML_SHIPPED)
    cfg, fl = Config(), Files()
    assert L.mcmcbh_read_namelist(str(p).encode(), C.byref(cfg), C.byref(fl)) == 0, L.mcmcbh_last_error()
    assert (cfg.method, cfg.nsimu, cfg.doadapt, cfg.adaptint, cfg.burnintime, cfg.doburnin) == (0, 1000, 1, 200, 1000, 1)
    assert (cfg.drscale, cfg.updatesigma, cfg.N0, cfg.S02) == (0.0, 1, 1.0, 0.0)
    assert (fl.chainfile, fl.ssfile, fl.s2file, fl.priorsfile) == (b"chain.dat", b"sschain.dat", b"s2chain.dat", b"")
    assert (fl.verbosity, fl.printint) == (1, 100)
    # everything not mentioned keeps MCMC_init_namelist's default (mcmcinit.F90:184-230) == the CUDA library's
    ref = mb.default_config()
    for k in ("adapthist", "adaptend", "initcmatn", "badaptint", "greedy", "scalelimit", "scalefactor", "condmax",
              "alphatarget", "nuparam"):
        assert getattr(cfg, k) == getattr(ref, k), k
    assert (fl.parfile, fl.cov0file, fl.sigma2file, fl.covffile, fl.meanfile) == (
        b"mcmcpar.dat", b"mcmccov.dat", b"mcmcsigma2.dat", b"mcmccovf.dat", b"mcmcmean.dat")


def test_namelist_forms_and_errors(tmp_path):
    L = hostlib()
    p = tmp_path / "a.nml"
    p.write_text("&MCMC nsimu=50, method=\"ram\", alphatarget = 0.3d0 nuparam=.66_dbl\n scalefactor = 2.0E0, "
                 "chainfile='my chain!.mat' ! comment / with slash\n/\n&mcmcb nchains = 4096, seed=7 model='expreg' /\n")
    cfg, fl = Config(), Files()
    assert L.mcmcbh_read_namelist(str(p).encode(), C.byref(cfg), C.byref(fl)) == 0, L.mcmcbh_last_error()
    assert (cfg.nsimu, cfg.method, cfg.alphatarget, cfg.nuparam, cfg.scalefactor) == (50, 1, 0.3, 0.66, 2.0)
    assert fl.chainfile == b"my chain!.mat" and (cfg.nchains, cfg.seed, cfg.model) == (4096, 7, b"expreg")
    for bad, code in (("&mcmc nsimu = 10, nosuchvar = 1 /", -2), ("&mcmc nsimu = ten /", -2), ("&mcmc nsimu = 10", -2),
                      ("&other x=1 /", -2)):
        p.write_text(bad)
        assert L.mcmcbh_read_namelist(str(p).encode(), C.byref(cfg), C.byref(fl)) == code
    assert L.mcmcbh_read_namelist(str(tmp_path / "missing.nml").encode(), C.byref(cfg), C.byref(fl)) == -1


def test_addtomat_streams_columns(tmp_path):
    # addtomat (matfiles.F90:187-293): the 'disk' save mode appends transpose(chain) blocks (MCMC_aux.F90:141-160)
    L = hostlib()
    rng = np.random.default_rng(1)
    blocks = [np.asfortranarray(rng.normal(size=(3, n))) for n in (4, 1, 6)]
    q = str(tmp_path / "chain.mat").encode()
    dp = C.POINTER(C.c_double)
    assert L.mcmcbh_write_mat4(q, b"chain", blocks[0].ctypes.data_as(dp), 3, 4, 3) == 0
    for b in blocks[1:]:
        assert L.mcmcbh_addto_mat4(q, b.ctypes.data_as(dp), 3, b.shape[1], 3) == 0
    assert np.array_equal(scipy.io.loadmat(q.decode())["chain"], np.hstack(blocks))
    bad = np.asfortranarray(rng.normal(size=(2, 2)))
    assert L.mcmcbh_addto_mat4(q, bad.ctypes.data_as(dp), 2, 2, 2) == -3          # "xmat should have same number of rows"
    assert L.mcmcbh_addto_mat4(str(tmp_path / "none.mat").encode(), bad.ctypes.data_as(dp), 2, 2, 2) == -1
    assert np.array_equal(scipy.io.loadmat(q.decode())["chain"], np.hstack(blocks))  # untouched by the failed calls


def test_namelist_write_back_round_trips(tmp_path):
    # write_mcmcinit_namelist (mcmcinit.F90:147-179, `nmlffile`): what is written reads back identically
    L = hostlib()
    p = tmp_path / "in.nml"
    p.write_text("&mcmc nsimu=777, method='er', drscale=1.5, adapthist=40, greedy=1, burnintime=30, doburnin=1, "
                 "condmax=1d10, chainfile='c h.mat', nmlffile='final.nml', S02=0.125, covnfile='n.dat', svddim=2, "
                 "condmaxini=1d12, sstype=5, sstrans=4.0 /\n"
                 "&mcmcb nchains=12, seed=99, model='gauss', datafile='d.dat', diag_stride=5, kernel=2 /\n")
    a, fa, b, fb = Config(), Files(), Config(), Files()
    assert L.mcmcbh_read_namelist(str(p).encode(), C.byref(a), C.byref(fa)) == 0, L.mcmcbh_last_error()
    out = tmp_path / "final.nml"
    assert L.mcmcbh_write_namelist(str(out).encode(), C.byref(a), C.byref(fa)) == 0, L.mcmcbh_last_error()
    assert L.mcmcbh_read_namelist(str(out).encode(), C.byref(b), C.byref(fb)) == 0, L.mcmcbh_last_error()
    assert bytes(a) == bytes(b) and bytes(fa) == bytes(fb)
    assert (b.method, b.nsimu, b.drscale, b.condmax, fb.chainfile, fb.nmlffile) == (3, 777, 1.5, 1e10, b"c h.mat", b"final.nml")
    # legacy variables are written back as read; sstype = 5 (t likelihood) switches the sigma2 update off (mcmcinit.F90:300-309)
    assert (fb.svddim, fb.sstype, fb.condmaxini, fb.sstrans, b.updatesigma) == (2, 5, 1e12, 4.0, 0)
    text = out.read_text()
    assert text.startswith("&mcmc") and "method = 'er'" in text and "&mcmcb" in text


def test_dat_and_mat4_writers(tmp_path):
    L = hostlib()
    rng = np.random.default_rng(0)
    x = np.asfortranarray(rng.normal(size=(7, 3)) * 10.0 ** rng.integers(-8, 8, size=(7, 3)))
    xp = x.ctypes.data_as(C.POINTER(C.c_double))
    p = str(tmp_path / "m.dat").encode()
    assert L.mcmcbh_write_dat(p, xp, 7, 3, 7) == 0
    assert np.array_equal(np.loadtxt(p.decode()), x)                    # %.17g round-trips exactly
    assert np.array_equal(mb.models.load_dat(p.decode()), x)
    data, r, c = C.POINTER(C.c_double)(), C.c_int(), C.c_int()
    assert L.mcmcbh_load_dat(p, C.byref(data), C.byref(r), C.byref(c)) == 0
    assert np.array_equal(np.ctypeslib.as_array(data, shape=(r.value, c.value)), x)
    L.mcmcbh_free(data)
    q = str(tmp_path / "m.mat").encode()
    assert L.mcmcbh_write_mat4(q, b"chain", xp, 7, 3, 7) == 0
    m = scipy.io.loadmat(q.decode())                                    # independent Level-1.0 (v4) reader
    assert np.array_equal(m["chain"], x)
    raw = open(q.decode(), "rb").read()                                  # header of matfiles.F90:41-48
    assert np.frombuffer(raw[:20], dtype="<i4").tolist() == [0, 7, 3, 0, 6] and raw[20:26] == b"chain\0"
    assert len(raw) == 20 + 6 + 7 * 3 * 8


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["shipped", "dram_mat"])
def test_driver_executable_writes_the_reference_files(tmp_path, variant):
    d = str(tmp_path)
    nml = NML_SHIPPED
    if variant == "dram_mat":
        nml = nml.replace("drscale     = 0", "drscale     = 2.0").replace("burnintime  = 1000", "burnintime  = 0") \
                 .replace("doburnin    = 1", "doburnin    = 0").replace("'chain.dat'", "'chain.mat'") \
                 .replace("adaptint    = 200", "adaptint    = 100\n initcmatn = 1")
    write_testcase(d, nml, "&mcmcb nchains = 3, seed = 42, store_chains = 2 /\n")
    r = subprocess.run([os.path.join(HOST, "mcmcb_main"), d], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    kw = dict(cases.NML_SHIPPED) if variant == "shipped" else dict(cases.NML_DRAM, nsimu=1000)
    blob = O.blob_expreg(cases.DATA_X, cases.DATA_Y)
    for c in range(2):
        ch = O.Chain(O.make_cfg(**kw), O.MODEL_EXPREG, blob, cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS)
        ch.philox(42, c)
        ch.run()
        ref = ch.results()
        suf = "" if c == 0 else "_%05d" % c
        if variant == "dram_mat":
            chain = scipy.io.loadmat(os.path.join(d, "chain%s.mat" % suf))["chain"]
        else:
            chain = np.loadtxt(os.path.join(d, "chain%s.dat" % suf), ndmin=2)
        ss = np.loadtxt(os.path.join(d, "sschain%s.dat" % suf), ndmin=2)
        s2 = np.loadtxt(os.path.join(d, "s2chain%s.dat" % suf), ndmin=2)
        assert chain.shape == ref["chain"].shape and np.array_equal(chain[:, -1], ref["chain"][:, -1])
        np.testing.assert_allclose(chain[:, :-1], ref["chain"][:, :-1], rtol=1e-10)
        np.testing.assert_allclose(ss[:, 0], ref["sschain"][:, 0], rtol=1e-9)
        np.testing.assert_allclose(s2[:, 0], ref["s2chain"][:, 0], rtol=1e-9)
        if c == 0:  # restart files, MCMC_aux.F90:46-66
            np.testing.assert_allclose(np.loadtxt(os.path.join(d, "mcmcparf.dat")), ref["par"], rtol=1e-10)
            np.testing.assert_allclose(np.loadtxt(os.path.join(d, "mcmcsigma2f.dat")), [ref["s2chain"][-1, 0], 11.0], rtol=1e-9)
            assert np.loadtxt(os.path.join(d, "mcmccovf.dat")).shape == (2, 2)
            assert np.loadtxt(os.path.join(d, "mcmcmean.dat")).shape == (2,)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["er", "scam"])
def test_driver_runs_the_other_samplers_of_the_testcase(tmp_path, method):
    # method = 'er' / 'scam' in mcmcinit.nml (mcmc_main.F90:29-37) through the driver executable, plus the namelist
    # write-back (nmlffile, MCMC_aux.F90:82-83)
    d = str(tmp_path)
    nml = NML_SHIPPED.replace("method = 'dram'", "method = '%s'\n nmlffile = 'final.nml'" % method) \
                     .replace("burnintime  = 1000", "burnintime  = 0").replace("doburnin    = 1", "doburnin    = 0") \
                     .replace("adaptint    = 200", "adaptint    = 100\n initcmatn = 2").replace("nsimu       = 1000", "nsimu       = 400")
    write_testcase(d, nml, "&mcmcb nchains = 2, seed = 7, store_chains = 1 /\n")
    r = subprocess.run([os.path.join(HOST, "mcmcb_main"), d], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    kw = dict(method=method, nsimu=400, doadapt=1, adaptint=100, initcmatn=2, burnintime=0, doburnin=0, drscale=0.0,
              updatesigma=1, N0=1.0, S02=0.0)
    ch = O.Chain(O.make_cfg(**kw), O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), cases.PAR0, cases.CMAT0,
                 cases.SIGMA2, cases.NOBS)
    ch.philox(7, 0)
    ch.run()
    ref = ch.results()
    chain = np.loadtxt(os.path.join(d, "chain.dat"), ndmin=2)
    assert chain.shape == ref["chain"].shape and np.array_equal(chain[:, -1], ref["chain"][:, -1])
    np.testing.assert_allclose(chain[:, :-1], ref["chain"][:, :-1], rtol=1e-8)
    text = open(os.path.join(d, "final.nml")).read()
    assert "method = '%s'" % method in text and "nsimu = 400" in text

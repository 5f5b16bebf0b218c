"""ctypes binding of the CPU oracle (oracle/mcmc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by mcmcf90_b200/.
Parity status (not pinned to a run of the Fortran; pinned by oracle/restate_np.py) -- see the header of mcmc_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DRAM, RAM, SCAM, ER = 0, 1, 2, 3
MODEL_EXPREG, MODEL_GAUSS, MODEL_BANANA, MODEL_HIER = 0, 1, 2, 3
METHODS = {"dram": DRAM, "am": DRAM, "ram": RAM, "scam": SCAM, "er": ER}


class Cfg(C.Structure):
    """Mirror of orc_cfg == namelist &mcmc (mcmcinit.F90:74-82)."""

    _fields_ = [
        ("method", C.c_int),
        ("nsimu", C.c_int), ("doadapt", C.c_int), ("adaptint", C.c_int), ("adapthist", C.c_int),
        ("adaptend", C.c_int), ("initcmatn", C.c_int),
        ("doburnin", C.c_int), ("burnintime", C.c_int), ("badaptint", C.c_int), ("greedy", C.c_int),
        ("scalelimit", C.c_double), ("scalefactor", C.c_double), ("drscale", C.c_double), ("condmax", C.c_double),
        ("N0", C.c_double), ("S02", C.c_double),
        ("updatesigma", C.c_int),
        ("alphatarget", C.c_double), ("nuparam", C.c_double),
        ("dodr", C.c_int), ("doscam", C.c_int), ("usesvd", C.c_int),
    ]


def build(force=False):
    so = os.path.join(_HERE, "libmcmcoracle.so")
    src = os.path.join(_HERE, "mcmc_oracle.c")
    hdr = os.path.join(_HERE, "mcmc_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmcmcoracle.so"], stdout=subprocess.DEVNULL)
    return so


_NATIVE = False


def use_native(on=True):
    """Timing only (bench.py's CPU baseline): switch to the -O3 -march=native build of the same source, compiled on
    the box it runs on.  The parity tests always use the reference-flag build."""
    global _NATIVE, _LIB
    if bool(on) != _NATIVE:
        _NATIVE, _LIB = bool(on), None


def _build_native():
    so = os.path.join(_HERE, "libmcmcoracle_native.so")
    src = os.path.join(_HERE, "mcmc_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmcmcoracle_native.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_build_native() if _NATIVE else build())
        dp = C.POINTER(C.c_double)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Cfg), C.c_int, dp, C.c_long, C.c_int, C.c_int, dp, dp, dp, C.POINTER(C.c_int)]
        L.orc_set_prior.argtypes = [C.c_void_p, dp, dp]
        L.orc_set_rng_injected.argtypes = [C.c_void_p, dp, C.c_long]
        L.orc_set_rng_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_run.argtypes = [C.c_void_p]
        L.orc_advance.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_pool.argtypes = [C.c_void_p, C.c_int]
        L.orc_factor_from_cov.argtypes = [C.c_void_p, dp]
        L.orc_set_R.argtypes = [C.c_void_p, dp]
        L.orc_free.argtypes = [C.c_void_p]
        for n in ("chain", "sschain", "s2chain", "R", "R2", "iC", "qcovstd", "cmat", "mean", "sigma2", "par"):
            f = getattr(L, "orc_%s_ptr" % n)
            f.restype = dp
            f.argtypes = [C.c_void_p]
        L.orc_wsum.restype = C.c_double
        L.orc_wsum.argtypes = [C.c_void_p]
        L.orc_erstayed.argtypes = [C.c_void_p]
        L.orc_counters.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.orc_default_cfg.argtypes = [C.POINTER(Cfg)]
        L.orc_check_params.argtypes = [C.POINTER(Cfg)]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_normals.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_gamma.restype = C.c_double
        L.orc_gamma.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_dtrmv_ut.argtypes = [C.c_int, dp, C.c_int, dp]
        L.orc_dgemv.argtypes = [C.c_char, C.c_int, dp, C.c_int, dp, dp]
        L.orc_dsymv_u.argtypes = [C.c_int, dp, C.c_int, dp, dp]
        L.orc_dpotf2_u.argtypes = [C.c_int, dp, C.c_int]
        L.orc_dpotri_u.argtypes = [C.c_int, dp, C.c_int]
        L.orc_drotg.argtypes = [dp, dp, dp, dp]
        L.orc_dchud.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp]
        L.orc_dchdd.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp]
        L.orc_covmat.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, C.c_int]
        L.orc_symeig.argtypes = [C.c_int, dp, dp, dp]
        L.orc_model_ss.restype = C.c_double
        L.orc_model_ss.argtypes = [C.c_int, dp, dp, C.c_int]
        L.orc_run_batch.argtypes = [C.POINTER(Cfg), C.c_int, dp, C.c_long, C.c_int, C.c_int, C.c_long, dp, dp, dp,
                                    C.POINTER(C.c_int), C.c_uint64, C.c_uint64, C.c_int, dp, dp, dp,
                                    C.POINTER(C.c_long), dp, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def make_cfg(**kw):
    """Namelist defaults (mcmcinit.F90:184-230) overridden by keywords; method may be a string."""
    c = Cfg()
    lib().orc_default_cfg(C.byref(c))
    for k, v in kw.items():
        if k == "method" and isinstance(v, str):
            v = METHODS[v.lower()]
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


# ---- model blobs (layout documented in include/mcmcb200_model.cuh) ----
def blob_expreg(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    npad = (n + 1) & ~1
    b = np.zeros(2 + 2 * npad)
    b[0] = n
    # +-max|x|, negative when some x is negative: lets the device model range-check once per evaluation and know
    # the sign of its exponents (mcmcb_expmul_direct)
    b[1] = (np.abs(x).max() if n else 0.0) * (-1.0 if n and x.min() < 0.0 else 1.0)
    b[2:2 + n] = x
    b[2 + npad:2 + npad + n] = y
    return b


def blob_gauss(mu, lam):
    mu = np.asarray(mu, dtype=np.float64)
    lam = np.asarray(lam, dtype=np.float64)
    d = mu.size
    dpad = (d + 1) & ~1
    b = np.zeros(2 + dpad + d * d)
    b[0] = d
    b[2:2 + d] = mu
    b[2 + dpad:] = lam.reshape(-1)
    return b


def blob_banana(d, bpar):
    return np.array([float(d), float(bpar)])


def blob_hier(y):
    y = np.asarray(y, dtype=np.float64)
    G, J = y.shape
    b = np.zeros(2 + G * J)
    b[0], b[1] = G, J
    b[2:] = y.T.reshape(-1)  # y[j][g]: observation j of every group contiguous
    return b


class Chain:
    """One reference-style run (one chain, module-global state in the reference: mcmc.F90:28-60)."""

    def __init__(self, cfg, model_id, blob, par0, cmat0, sigma2, nobs, prior=None):
        L = lib()
        self.cfg = cfg
        self.npar = int(len(par0))
        self.blob = np.ascontiguousarray(blob, dtype=np.float64)
        self.par0 = np.ascontiguousarray(par0, dtype=np.float64)
        self.cmat0 = np.asfortranarray(np.asarray(cmat0, dtype=np.float64))
        self.sigma2_0 = np.ascontiguousarray(np.atleast_1d(sigma2), dtype=np.float64)
        self.nycol = int(self.sigma2_0.size)
        self.nobs = np.ascontiguousarray(np.atleast_1d(nobs), dtype=np.int32)
        self.h = L.orc_create(C.byref(cfg), model_id, _dp(self.blob), self.blob.size, self.npar, self.nycol,
                              _dp(self.par0), _dp(self.cmat0), _dp(self.sigma2_0),
                              self.nobs.ctypes.data_as(C.POINTER(C.c_int)))
        self._keep = []
        if prior is not None:
            mu = np.ascontiguousarray(prior[0], dtype=np.float64)
            sg = np.ascontiguousarray(prior[1], dtype=np.float64)
            L.orc_set_prior(self.h, _dp(mu), _dp(sg))

    def inject(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._keep.append(u)
        lib().orc_set_rng_injected(self.h, _dp(u), u.size)

    def philox(self, seed, chain_id):
        lib().orc_set_rng_philox(self.h, seed, chain_id)

    def run(self):
        return lib().orc_run(self.h)

    def advance(self, upto):
        """Run up to step index `upto` (simuind); resumable."""
        return lib().orc_advance(self.h, int(upto))

    def set_pool(self, on=True):
        lib().orc_set_pool(self.h, 1 if on else 0)

    def factor_from_cov(self, cov):
        cov = np.asfortranarray(np.asarray(cov, dtype=np.float64))
        return lib().orc_factor_from_cov(self.h, _dp(cov))

    def set_R(self, R):
        R = np.asfortranarray(np.asarray(R, dtype=np.float64))
        lib().orc_set_R(self.h, _dp(R))

    def _arr(self, name, shape, order="F"):
        p = getattr(lib(), "orc_%s_ptr" % name)(self.h)
        n = int(np.prod(shape))
        a = np.ctypeslib.as_array(p, shape=(n,)).copy()
        return a.reshape(shape, order=order)

    def counters(self):
        out = (C.c_long * 8)()
        lib().orc_counters(self.h, out)
        k = ["stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "status", "ndrawn"]
        r = dict(zip(k, [int(v) for v in out]))
        r["erstayed"] = int(lib().orc_erstayed(self.h))
        return r

    def results(self):
        ns, d, m = self.cfg.nsimu, self.npar, self.nycol
        cnt = self.counters()
        r = dict(cnt)
        r["chain"] = self._arr("chain", (ns, d + 1))[:cnt["chainind"]]
        r["sschain"] = self._arr("sschain", (ns, m + 1))[:cnt["chainind"]]
        r["s2chain"] = self._arr("s2chain", (ns, m))
        r["R"] = self._arr("R", (d, d))
        r["R2"] = self._arr("R2", (d, d))
        r["iC"] = self._arr("iC", (d, d))
        r["qcovstd"] = self._arr("qcovstd", (d,))
        r["cmat"] = self._arr("cmat", (d, d))
        r["mean"] = self._arr("mean", (d,))
        r["sigma2"] = self._arr("sigma2", (m,))
        r["par"] = self._arr("par", (d,))
        r["wsum"] = lib().orc_wsum(self.h)
        return r

    def close(self):
        if self.h:
            lib().orc_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_batch(cfg, model_id, blob, par0, cmat0, sigma2, nobs, seed=0, chain0=0, nthreads=1, moments=False):
    """Run par0.shape[0] independent chains, one per host thread. Returns dict of arrays + seconds."""
    L = lib()
    blob = np.ascontiguousarray(blob, dtype=np.float64)
    par0 = np.ascontiguousarray(par0, dtype=np.float64)
    N, d = par0.shape
    cmat0 = np.asfortranarray(np.asarray(cmat0, dtype=np.float64))
    sigma2 = np.ascontiguousarray(np.atleast_1d(sigma2), dtype=np.float64)
    nobs = np.ascontiguousarray(np.atleast_1d(nobs), dtype=np.int32)
    last = np.zeros((N, d))
    mean = np.zeros((N, d))
    cm = np.zeros((N, d, d))
    cnt = np.zeros((N, 8), dtype=np.int64)
    cmean = np.zeros((N, d)) if moments else None
    ccov = np.zeros((N, d, d)) if moments else None
    sec = C.c_double(0.0)
    L.orc_run_batch(C.byref(cfg), model_id, _dp(blob), blob.size, d, sigma2.size, N, _dp(par0), _dp(cmat0),
                    _dp(sigma2), nobs.ctypes.data_as(C.POINTER(C.c_int)), seed, chain0, nthreads,
                    _dp(last), _dp(mean), _dp(cm), cnt.ctypes.data_as(C.POINTER(C.c_long)),
                    _dp(cmean), _dp(ccov), C.byref(sec))
    return dict(par=last, mean=mean, cmat=cm, counters=cnt, chain_mean=cmean, chain_cov=ccov, seconds=sec.value)


def pooled_moments(wsum, mean, cmat):
    """Merge of per-chain accumulators (wsum (N,), mean (N,d), cmat (N,d,d)) the way
    mcmcf90_b200/csrc/pool.cuh defines it: W, mu, cov = S2 / (W - 1)."""
    w = np.asarray(wsum, dtype=np.float64)
    use = w > 0
    w, mean, cmat = w[use], np.asarray(mean)[use], np.asarray(cmat)[use]
    W = w.sum()
    mu = (w[:, None] * mean).sum(0) / W
    dm = mean - mu
    S2 = ((w - 1.0)[:, None, None] * cmat + w[:, None, None] * dm[:, :, None] * dm[:, None, :]).sum(0)
    return W, mu, S2 / (W - 1.0)


def run_pooled(cfg, model_id, blob, par0, cmat0, sigma2, nobs, seed=0, chain0=0, prior=None):
    """Reference for pool_adapt = 1: N chains advanced in lockstep; at every pooled tick (the AM
    branch of MCMC_adapt for DRAM/AM/SCAM, every adaptint steps for RAM) the accumulators are
    merged and every chain takes the factor of the pooled covariance.  Returns the Chain objects and
    the list of (step, W, mu, cov) ticks."""
    par0 = np.asarray(par0, dtype=np.float64)
    N, d = par0.shape
    chains = []
    for c in range(N):
        ch = Chain(cfg, model_id, blob, par0[c], cmat0, sigma2, nobs, prior=prior)
        ch.philox(seed, chain0 + c)
        ch.set_pool(True)
        chains.append(ch)
    cc = chains[0].cfg
    lib().orc_check_params(C.byref(cc))
    ram = cc.method == RAM
    ticks = []
    i = 1
    while i < cc.nsimu:
        nxt = min(cc.nsimu, (i // cc.adaptint + 1) * cc.adaptint)
        for ch in chains:
            ch.advance(nxt)
        i = nxt
        is_tick = i % cc.adaptint == 0 and not (cc.adaptend > 0 and i > cc.adaptend)
        if not ram:
            is_tick = is_tick and i >= cc.burnintime + cc.adaptint + cc.adapthist
        if not is_tick:
            continue
        if ram:
            Rs = np.array([ch._arr("R", (d, d)) for ch in chains])
            Rs = np.triu(Rs)
            S = np.einsum("cki,ckj->ij", Rs, Rs) / N
            R = np.linalg.cholesky(S).T
            for ch in chains:
                ch.set_R(R)
            ticks.append((i, float(N), None, S))
        else:
            w = np.array([lib().orc_wsum(ch.h) for ch in chains])
            m = np.array([ch._arr("mean", (d,)) for ch in chains])
            cm = np.array([ch._arr("cmat", (d, d)) for ch in chains])
            cm = np.triu(cm) + np.transpose(np.triu(cm, 1), (0, 2, 1))
            W, mu, cov = pooled_moments(w, m, cm)
            for ch in chains:
                ch.factor_from_cov(cov)
            ticks.append((i, W, mu, cov))
    return chains, ticks

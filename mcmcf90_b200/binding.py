"""ctypes binding of include/mcmcb200.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DRAM, RAM, SCAM, ER = 0, 1, 2, 3
RNG_PHILOX, RNG_INJECTED = 0, 1
METHODS = {"dram": DRAM, "am": DRAM, "ram": RAM, "scam": SCAM, "er": ER}
ERRORS = {-1: "EINVAL", -2: "ECUDA", -3: "EUNSUPPORTED", -4: "ENOMODEL", -5: "ENOMEM"}
COUNTER_NAMES = ["stayed", "bndstayed", "draccepted", "drtries", "chainind", "simuind", "status", "ndrawn"]

EXPORTS = ["mcmcb_default_config", "mcmcb_check_config", "mcmcb_create", "mcmcb_destroy", "mcmcb_last_error",
           "mcmcb_set_data", "mcmcb_set_priors", "mcmcb_set_initial", "mcmcb_inject_uniforms", "mcmcb_run",
           "mcmcb_sync", "mcmcb_fetch_chain", "mcmcb_fetch", "mcmcb_dump_pop", "mcmcb_stream",
           "mcmcb_launch_count", "mcmcb_info", "mcmcb_chains_per_thread", "mcmcb_dfma_peak", "mcmcb_exp_selftest",
           "mcmcb_set_allreduce", "mcmcb_pool_fetch", "mcmcb_diagnostics", "mcmcb_diag_reset", "mcmcb_load_plugin",
           "mcmcb_dump_pop_ex", "mcmcb_stream_of", "mcmcb_ngpus", "mcmcb_nccl_calls", "mcmcb_fetch_stats"]

# int fn(void* user, double* device_buf, size_t n, void* cuda_stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class MCMCBError(RuntimeError):
    pass


class Config(C.Structure):
    """mcmcb_config: namelist &mcmc (mcmcinit.F90:74-82) + batch fields."""
    _fields_ = [
        ("abi_version", C.c_int),
        ("method", C.c_int), ("nsimu", C.c_int),
        ("doadapt", C.c_int), ("adaptint", C.c_int), ("adapthist", C.c_int), ("adaptend", C.c_int),
        ("initcmatn", C.c_int),
        ("doburnin", C.c_int), ("burnintime", C.c_int), ("badaptint", C.c_int), ("greedy", C.c_int),
        ("scalelimit", C.c_double), ("scalefactor", C.c_double), ("drscale", C.c_double), ("condmax", C.c_double),
        ("N0", C.c_double), ("S02", C.c_double),
        ("updatesigma", C.c_int),
        ("alphatarget", C.c_double), ("nuparam", C.c_double),
        ("nchains", C.c_longlong), ("chain_offset", C.c_longlong), ("seed", C.c_ulonglong),
        ("rng_mode", C.c_int), ("device", C.c_int), ("store_chains", C.c_int), ("lanes_per_chain", C.c_int),
        ("dump_stride", C.c_int), ("kernel", C.c_int),
        ("pool_adapt", C.c_int), ("diag_stride", C.c_int), ("diag_lags", C.c_int), ("ngpus", C.c_int),
        ("model", C.c_char * 32),
    ]


def library_path():
    return os.path.join(_HERE, "libmcmcb200.so")


def load_library():
    """Load the CUDA library; raise (never fall back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise MCMCBError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.mcmcb_default_config.argtypes = [C.POINTER(Config)]
    L.mcmcb_check_config.argtypes = [C.POINTER(Config), ip, ip, ip]
    L.mcmcb_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.mcmcb_destroy.argtypes = [C.c_void_p]
    L.mcmcb_last_error.argtypes = [C.c_void_p]
    L.mcmcb_last_error.restype = C.c_char_p
    L.mcmcb_set_data.argtypes = [C.c_void_p, dp, C.c_size_t]
    L.mcmcb_set_priors.argtypes = [C.c_void_p, dp, dp, C.c_int]
    L.mcmcb_set_initial.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_longlong, dp, dp, ip]
    L.mcmcb_inject_uniforms.argtypes = [C.c_void_p, dp, C.c_size_t]
    L.mcmcb_run.argtypes = [C.c_void_p, C.c_int]
    L.mcmcb_sync.argtypes = [C.c_void_p]
    L.mcmcb_fetch_chain.argtypes = [C.c_void_p, C.c_longlong, C.c_int, dp, dp, dp, ip]
    L.mcmcb_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
    L.mcmcb_dump_pop.argtypes = [C.c_void_p, dp, C.c_size_t, ip]
    L.mcmcb_fetch_stats.argtypes = [C.c_void_p, C.c_longlong, dp, dp, dp, dp, dp, C.POINTER(C.c_longlong)]
    L.mcmcb_dump_pop_ex.argtypes = [C.c_void_p, dp, dp, dp, C.c_size_t, ip]
    L.mcmcb_stream.argtypes = [C.c_void_p]
    L.mcmcb_stream.restype = C.c_void_p
    L.mcmcb_stream_of.argtypes = [C.c_void_p, C.c_int]
    L.mcmcb_stream_of.restype = C.c_void_p
    L.mcmcb_ngpus.argtypes = [C.c_void_p]
    L.mcmcb_nccl_calls.argtypes = [C.c_void_p]
    L.mcmcb_nccl_calls.restype = C.c_longlong
    L.mcmcb_launch_count.argtypes = [C.c_void_p]
    L.mcmcb_launch_count.restype = C.c_longlong
    L.mcmcb_info.argtypes = [C.c_void_p, ip, ip, ip, ip, ip, ip, C.POINTER(C.c_size_t)]
    L.mcmcb_chains_per_thread.argtypes = [C.c_void_p]
    L.mcmcb_dfma_peak.argtypes = [C.c_int, dp, dp]
    L.mcmcb_exp_selftest.argtypes = [C.c_int, dp, C.c_double, dp, dp, C.c_size_t]
    L.mcmcb_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p]
    L.mcmcb_pool_fetch.argtypes = [C.c_void_p, dp, dp, dp]
    L.mcmcb_diagnostics.argtypes = [C.c_void_p, dp, dp, dp, dp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.mcmcb_diag_reset.argtypes = [C.c_void_p]
    L.mcmcb_load_plugin.argtypes = [C.c_char_p]
    _LIB = L
    return L


def default_config(**kw):
    """Namelist defaults (MCMC_init_namelist, mcmcinit.F90:184-230) overridden by keywords."""
    c = Config()
    load_library().mcmcb_default_config(C.byref(c))
    for k, v in kw.items():
        if k == "method" and isinstance(v, str):
            v = METHODS[v.lower()]
        if k == "model" and isinstance(v, str):
            v = v.encode()
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


def load_plugin(path):
    """Load a user-model plugin library (include/mcmcb200_plugin.cuh); its models become available by name."""
    rc = load_library().mcmcb_load_plugin(os.fspath(path).encode())
    if rc:
        raise MCMCBError("mcmcb_load_plugin(%s): %s" % (path, ERRORS.get(rc, rc)))


def dfma_peak(device=0):
    t, ms = C.c_double(0), C.c_double(0)
    rc = load_library().mcmcb_dfma_peak(device, C.byref(t), C.byref(ms))
    if rc:
        raise MCMCBError("mcmcb_dfma_peak failed: %s" % ERRORS.get(rc, rc))
    return t.value, ms.value


def exp_selftest(a, scale=1.0, device=0):
    """(mcmcb_exp_fast(a), mcmcb_expmul_fast(a, scale)) evaluated on the device."""
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    o1, o2 = np.empty_like(a), np.empty_like(a)
    rc = load_library().mcmcb_exp_selftest(device, _dp(a), float(scale), _dp(o1), _dp(o2), a.size)
    if rc:
        raise MCMCBError("mcmcb_exp_selftest failed: %s" % ERRORS.get(rc, rc))
    return o1, o2


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Sampler:
    """One handle = nchains independent chains on cfg.ngpus GPUs (replaces the module-global
    single chain of mcmc.F90:28-60)."""

    def __init__(self, cfg):
        self.L = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        self._chk(self.L.mcmcb_create(C.byref(cfg), C.byref(self.h)), "mcmcb_create")
        self.nchains = int(cfg.nchains)
        self.npar = self.nycol = None
        self._keep = []

    def _chk(self, rc, what):
        if rc < 0:
            msg = self.L.mcmcb_last_error(self.h).decode() if self.h else ""
            raise MCMCBError("%s: %s %s" % (what, ERRORS.get(rc, rc), msg))
        return rc

    def set_data(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.float64)
        self._chk(self.L.mcmcb_set_data(self.h, _dp(blob), blob.size), "mcmcb_set_data")

    def set_priors(self, mu, sig):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        self._chk(self.L.mcmcb_set_priors(self.h, _dp(mu), _dp(sig), mu.size), "mcmcb_set_priors")

    def set_initial(self, par0, cmat0, sigma2, nobs):
        """par0: (npar,) shared or (nchains, npar); cmat0 (npar,npar); sigma2, nobs (nycol,)."""
        par0 = np.ascontiguousarray(par0, dtype=np.float64)
        cmat0 = np.asfortranarray(np.asarray(cmat0, dtype=np.float64))
        sigma2 = np.ascontiguousarray(np.atleast_1d(sigma2), dtype=np.float64)
        nobs = np.ascontiguousarray(np.atleast_1d(nobs), dtype=np.int32)
        npar = par0.shape[-1]
        stride = 0 if par0.ndim == 1 else npar
        if par0.ndim == 2 and par0.shape[0] != self.nchains:
            raise ValueError("par0 must have nchains rows")
        self.npar, self.nycol = int(npar), int(sigma2.size)
        self._chk(self.L.mcmcb_set_initial(self.h, npar, sigma2.size, _dp(par0), stride, _dp(cmat0), _dp(sigma2),
                                           nobs.ctypes.data_as(C.POINTER(C.c_int))), "mcmcb_set_initial")

    def inject_uniforms(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        if u.ndim == 1:
            u = u.reshape(1, -1)
        if u.shape[0] != self.nchains:
            raise ValueError("u must have nchains rows")
        self._chk(self.L.mcmcb_inject_uniforms(self.h, _dp(u), u.shape[1]), "mcmcb_inject_uniforms")

    def run(self, nsteps, sync=True):
        self._chk(self.L.mcmcb_run(self.h, int(nsteps)), "mcmcb_run")
        if sync:
            self.sync()

    def sync(self):
        self._chk(self.L.mcmcb_sync(self.h), "mcmcb_sync")

    def fetch(self, what, out=None):
        """Per-chain state, chain-major.  `out` (optional) is a C-contiguous array of the right size to
        receive the copy -- pass a view of pinned host memory to get an asynchronous-speed transfer."""
        d, m, N = self.npar, self.nycol, self.nchains
        if what == "counters":
            shape, dt = (N, 8), np.int64
        elif what == "erstayed":
            shape, dt = (N,), np.int64
        else:
            width = {"par": d, "ss": m, "sspri": 1, "sigma2": m, "mean": d, "wsum": 1, "qcovstd": d,
                     "cmat": d * d, "R": d * d, "R2": d * d, "iC": d * d}[what]
            shape, dt = (N, width), np.float64
        if out is None:
            out = np.empty(shape, dtype=dt)
        else:
            if out.dtype != dt or out.size != shape[0] * shape[1] or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous %s array of %d elements" % (dt.__name__, shape[0] * shape[1]))
            out = out.reshape(shape)
        self._chk(self.L.mcmcb_fetch(self.h, what.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes),
                  "mcmcb_fetch(%s)" % what)
        if what in ("cmat", "R", "R2", "iC"):
            out = out.reshape(N, d, d).transpose(0, 2, 1)  # column-major -> [chain, i, j]
        return out

    def counters(self, out=None):
        c = self.fetch("counters", out=out)
        return {k: c[:, i] for i, k in enumerate(COUNTER_NAMES)}

    def fetch_chain(self, chain):
        """Returns dict(chain=(rows, npar+1), sschain=(rows, nycol+1), s2chain=(simuind, nycol))."""
        ld = int(self.cfg.nsimu)
        d, m = self.npar, self.nycol
        ch = np.zeros((ld, d + 1), order="F")
        ss = np.zeros((ld, m + 1), order="F")
        s2 = np.zeros((ld, m), order="F")
        nrows = C.c_int(0)
        self._chk(self.L.mcmcb_fetch_chain(self.h, chain, ld, _dp(ch), _dp(ss), _dp(s2), C.byref(nrows)),
                  "mcmcb_fetch_chain")
        n = nrows.value
        return dict(chain=ch[:n].copy(), sschain=ss[:n].copy(), s2chain=s2.copy(), nrows=n)

    def fetch_stats(self, chain):
        """One chain's (mean, cmat, wsum, R, sigma2, counters) through the typed entry point mcmcb_fetch_stats."""
        d, m = self.npar, self.nycol
        mean, cm, R, s2 = np.zeros(d), np.zeros((d, d), order="F"), np.zeros((d, d), order="F"), np.zeros(m)
        w, cnt = C.c_double(0), (C.c_longlong * 8)()
        self._chk(self.L.mcmcb_fetch_stats(self.h, int(chain), _dp(mean), _dp(cm), C.byref(w), _dp(R), _dp(s2), cnt), "mcmcb_fetch_stats")
        return dict(mean=mean, cmat=np.ascontiguousarray(cm), wsum=w.value, R=np.ascontiguousarray(R), sigma2=s2,
                    counters=dict(zip(COUNTER_NAMES, [int(v) for v in cnt])))

    def set_allreduce(self, fn):
        """fn(device_ptr:int, n:int, cuda_stream:int) -> None sum-reduces n doubles in place over all
        ranks (mcmcf90_b200.parallel.attach builds it from torch.distributed); None detaches."""
        if fn is None:
            self._ar = ALLREDUCE_FN(0)
        else:
            def _cb(user, ptr, n, stream):
                try:
                    fn(int(ptr), int(n), int(stream or 0))
                    return 0
                except Exception as e:  # the C side turns this into an error code
                    self._ar_error = e
                    return 1
            self._ar = ALLREDUCE_FN(_cb)
        self._chk(self.L.mcmcb_set_allreduce(self.h, self._ar, None), "mcmcb_set_allreduce")

    def pool_fetch(self):
        """(wsum, mean[npar], cov[npar, npar]) of the last pooled adaptation tick."""
        d = self.npar
        w, m, cv = C.c_double(0), np.zeros(d), np.zeros((d, d), order="F")
        self._chk(self.L.mcmcb_pool_fetch(self.h, C.byref(w), _dp(m), _dp(cv)), "mcmcb_pool_fetch")
        return w.value, m, np.ascontiguousarray(cv)

    def diagnostics(self):
        """dict(rhat, ess, mean, var (npar,), nsnap, nchains): collective over the attached ranks."""
        d = self.npar
        r, e, m, v = np.zeros(d), np.zeros(d), np.zeros(d), np.zeros(d)
        ns, nc = C.c_longlong(0), C.c_longlong(0)
        self._chk(self.L.mcmcb_diagnostics(self.h, _dp(r), _dp(e), _dp(m), _dp(v), C.byref(ns), C.byref(nc)),
                  "mcmcb_diagnostics")
        return dict(rhat=r, ess=e, mean=m, var=v, nsnap=ns.value, nchains=nc.value)

    def diag_reset(self):
        self._chk(self.L.mcmcb_diag_reset(self.h), "mcmcb_diag_reset")

    def dump_pop(self):
        out = np.zeros((self.nchains, self.npar))
        step = C.c_int(0)
        rc = self._chk(self.L.mcmcb_dump_pop(self.h, _dp(out), out.nbytes, C.byref(step)), "mcmcb_dump_pop")
        return (step.value, out) if rc == 1 else None

    def dump_pop_ex(self):
        """(step, theta (nchains, npar), ss (nchains, nycol), sigma2 (nchains, nycol)) of the oldest finished snapshot."""
        par = np.zeros((self.nchains, self.npar))
        ss, s2 = np.zeros((self.nchains, self.nycol)), np.zeros((self.nchains, self.nycol))
        step = C.c_int(0)
        rc = self._chk(self.L.mcmcb_dump_pop_ex(self.h, _dp(par), _dp(ss), _dp(s2), par.nbytes, C.byref(step)), "mcmcb_dump_pop_ex")
        return (step.value, par, ss, s2) if rc == 1 else None

    @property
    def ngpus(self):
        return int(self.L.mcmcb_ngpus(self.h))

    @property
    def nccl_calls(self):
        return int(self.L.mcmcb_nccl_calls(self.h))

    def stream_of(self, k):
        return self.L.mcmcb_stream_of(self.h, int(k))

    @property
    def stream(self):
        return self.L.mcmcb_stream(self.h)

    @property
    def launches(self):
        return int(self.L.mcmcb_launch_count(self.h))

    def info(self):
        v = [C.c_int(0) for _ in range(6)]
        sm = C.c_size_t(0)
        self.L.mcmcb_info(self.h, *[C.byref(x) for x in v], C.byref(sm))
        k = ["npar", "nycol", "lanes_per_chain", "kernel", "threads_per_block", "blocks"]
        r = dict(zip(k, [x.value for x in v]))
        r["smem_bytes"] = sm.value
        r["chains_per_thread"] = int(self.L.mcmcb_chains_per_thread(self.h))
        return r

    def close(self):
        if self.h:
            self.L.mcmcb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

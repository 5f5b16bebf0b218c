"""restate_np.py -- SECOND, independent CPU restatement of mcmcf90's sampling loops, in numpy.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ to pin the C oracle
(mcmc_oracle.c).  Never imported by mcmcf90_b200/, never timed, never shipped.

Why it exists.  Neither this container nor the GPU box has a Fortran compiler (probed: gfortran, f95, flang,
nvfortran, ifort, lfortran, f951 -- all absent; profiles/r02_probe.txt), so the reference cannot be run, and it
ships no expected outputs.  The C oracle is therefore checked against a second reading of the Fortran that shares
NO code with it and differs from it in every third-party routine:

  * written straight from the .F90/.f sources (citations are file:line into /root/reference), array-at-a-time the
    way the Fortran is (whole `chain(nsimu,npar+1)` array, `covmat` on array sections, `lastind/lastfreq` `save`
    variables), not from mcmc_oracle.c;
  * BLAS/LAPACK calls go to the REAL libraries through scipy (OpenBLAS: dtrmv, dgemv, dsymv, dpotrf, dpotri, dgesvd,
    drotg, ddot, dnrm2) -- what a user linking `-llapack -lblas` (testcases/Makefile:11) gets -- where the C oracle
    restates netlib loops;
  * libm through numpy.

Agreement of the two on the same injected uniform stream (tests/test_ref_parity.py: counters and chain indices
exact, values to 1e-10 relative) means a misreading would have to be made twice, independently, in the same way.
The recipe that builds the actual Fortran with its `random_number` calls redirected to the same stream is
oracle/_ref/Makefile; it runs wherever gfortran exists.

dgesvd returns singular vectors with implementation-defined signs; the sign of column j decides the direction of a
SCAM move.  To make SVD-based runs comparable, `fix_signs` applies the convention the C oracle documents (largest-
magnitude component of each column positive) -- a declared normalisation, not part of the reference.
"""
import numpy as np
from scipy.linalg import blas, lapack

TINY = np.finfo(np.float64).tiny
LOG_REALMIN = float(np.log(TINY))          # mcmcprec.F90:34-41
HUGE = float(np.finfo(np.float64).max)     # huge(0.0d0), MCMC_run.F90:50


class Stream:
    """Replaces the compiler runtime's random_number (mcmcrand.F90:55,104,138,156,177; MCMC_DRAM.F90:132,151): the
    next n numbers of an injected stream."""

    def __init__(self, u):
        self.u = np.asarray(u, dtype=np.float64)
        self.n = 0
        # normal_bm's `save`d spare, mcmcrand.F90:172-173
        self.saved = False
        self.saved_y = 0.0

    def random_number(self, n=1):
        if self.n + n > self.u.size:
            raise RuntimeError("injected stream exhausted")
        r = self.u[self.n:self.n + n]
        self.n += n
        return r.copy()

    # mcmcrand.F90:166-190
    def normal_bm(self):
        if not self.saved:
            while True:
                x = self.random_number(2)
                x = 2.0 * x - 1.0
                xx = x[0] ** 2 + x[1] ** 2
                if xx < 1.0 and xx != 0.0:
                    break
            z = np.sqrt(-2.0 * np.log(xx) / xx)
            self.saved_y = z * x[0]
            self.saved = True
            return z * x[1]
        self.saved = False
        return self.saved_y

    # mcmcrand.F90:60-83
    def random_normal(self, n):
        return np.array([self.normal_bm() * 1.0 + 0.0 for _ in range(n)])

    # mcmcrand.F90:120-162
    def gammar_mt(self, a, b):
        aa, bb = a, b
        if aa < 1.0:
            u = self.random_number(1)[0]
            bb = bb * u ** (1.0 / aa)
            aa = aa + 1.0
        d = aa - 1.0 / 3.0
        c = 1.0 / np.sqrt(9.0 * d)
        while True:
            while True:
                x = self.random_normal(1)[0]
                v = 1.0 + c * x
                if v > 0.0:
                    break
            v = v ** 3
            u = self.random_number(1)[0]
            if u < 1.0 - 0.0331 * x ** 4:
                break
            if np.log(u) < 0.5 * x ** 2 + d * (1.0 - v + np.log(v)):
                break
        return bb * d * v

    # mcmcrand.F90:86-111, n = 1
    def random_gamma(self, a, b):
        if a < 1.0:
            u = self.random_number(1)[0]
            return self.gammar_mt(1.0 + a, b) * u ** (1.0 / a)
        return self.gammar_mt(a, b)


# ------------------------------------------------------------------------------------------------ matutils.F90
def matmulu_t(x, v):
    """matmulu(x,v,'t'), matutils.F90:84-110: dcopy + dtrmv('u','t','n')."""
    return blas.dtrmv(np.asfortranarray(x), v.copy(), lower=0, trans=1, diag=0)


def matmulx(x, v, trans="n"):
    """matutils.F90:137-162: dgemv."""
    return blas.dgemv(1.0, np.asfortranarray(x), v, trans=1 if trans in "tT" else 0)


def matmuls(x, v):
    """matutils.F90:167-181: dsymv('u')."""
    return blas.dsymv(1.0, np.asfortranarray(x), v, lower=0)


def covmat(x, cmat, w, xmean, wsum, update):
    """matutils.F90:232-341.  x is rows x p; w has one entry per row or a single entry; returns (cmat, xmean, wsum)."""
    n, p = x.shape
    w = np.atleast_1d(np.asarray(w, dtype=np.float64))
    if w.size == n:                      # :260-262  (a one-row window with one weight also lands here)
        w2, wsum2 = -1.0, w.sum()
    elif w.size == 1:                    # :263-265
        w2, wsum2 = w[0], n * w[0]
    else:
        raise ValueError("covmat: invalid weight in w")
    doupdate = bool(update) and wsum > 0.0   # :274-283
    if doupdate:
        cmat = cmat.copy()
        xmean = xmean.copy()
        for i in range(n):               # :287-310
            xmean2 = x[i, :] - xmean
            w3 = w[i] if w2 == -1.0 else w2
            cmat = cmat + w3 / (wsum + w3 - 1.0) * (wsum / (wsum + w3) * np.outer(xmean2, xmean2) - cmat)
            xmean = xmean + w3 / (wsum + w3) * xmean2
            wsum = w3 + wsum
        return cmat, xmean, wsum
    xmean2 = np.empty(p)                 # :312-337
    for i in range(p):
        xmean2[i] = np.sum(x[:, i] * (w if w2 == -1.0 else w2)) / wsum2
    cmat = np.array(cmat, dtype=np.float64, copy=True)
    for i in range(p):
        for j in range(i + 1):
            cmat[i, j] = np.dot(x[:, i] - xmean2[i], (x[:, j] - xmean2[j]) * (w if w2 == -1.0 else w2)) / (wsum2 - 1.0)
            if i != j:
                cmat[j, i] = cmat[i, j]
    return cmat, xmean2, wsum2


def fix_signs(u):
    """Declared normalisation of dgesvd's arbitrary column signs (module docstring)."""
    u = u.copy()
    for j in range(u.shape[1]):
        k = int(np.argmax(np.abs(u[:, j])))
        if u[k, j] < 0.0:
            u[:, j] = -u[:, j]
    return u


def _svd(cmat):
    """dgesvd('A','N',...) on a copy of cmat, matutils.F90:409,615."""
    u, s, _vt, info = lapack.dgesvd(np.asfortranarray(cmat.copy()), compute_uv=1, full_matrices=1)
    return fix_signs(u), s.copy(), info


def covtor_svd(cmat, condmax):
    """matutils.F90:378-453.  Returns (R, info)."""
    n = cmat.shape[0]
    u, s, info2 = _svd(cmat)
    if s[0] == 0.0:
        return None, n
    tol = s[0] / condmax                 # :421-423 (the earlier assignments are dead)
    if s[n - 1] <= tol:
        s = np.where(s < tol, tol, s)
        info2 = -1
    R = u.copy()
    for i in range(n):
        R[:, i] = np.sqrt(s[i]) * R[:, i]   # dscal, :442
    return R, info2


def scam_svd(cmat, condmax):
    """matutils.F90:583-653.  Returns (R, std, info)."""
    n = cmat.shape[0]
    u, s, info2 = _svd(cmat)
    if s[0] == 0.0:
        return None, None, n
    tol = s[0] / condmax
    if s[n - 1] <= tol:
        s = np.where(s < tol, tol, s)
        info2 = -1
    return u.copy(), np.sqrt(s), info2


def dchud(r, x):
    """dchud.f:122-139 (nz = 0, matutils.F90:680-682): r is upper triangular, updated in place."""
    p = r.shape[0]
    c = np.zeros(p)
    s = np.zeros(p)
    for j in range(p):
        xj = x[j]
        for i in range(j):
            t = c[i] * r[i, j] + s[i] * xj
            xj = c[i] * xj - s[i] * r[i, j]
            r[i, j] = t
        rr, _z, c[j], s[j] = _drotg(r[j, j], xj)
        r[j, j] = rr


def _drotg(a, b):
    """BLAS drotg with its in/out arguments: returns (r, z, c, s).  scipy exposes only (c, s), so r is recomputed the
    way every drotg defines it: r = c*a + s*b."""
    c, s = blas.drotg(a, b)
    return c * a + s * b, 0.0, c, s


def dchdd(r, x):
    """dchdd.f:141-179 (nz = 0): returns info (0 or -1); r updated in place unless info = -1."""
    p = r.shape[0]
    s = np.zeros(p)
    c = np.zeros(p)
    s[0] = x[0] / r[0, 0]
    for j in range(1, p):
        s[j] = x[j] - blas.ddot(r[:j, j], s[:j])
        s[j] = s[j] / r[j, j]
    norm = blas.dnrm2(s)
    if not norm < 1.0:
        return -1
    alpha = np.sqrt(1.0 - norm ** 2)
    for ii in range(1, p + 1):
        i = p - ii
        scale = alpha + abs(s[i])
        a = alpha / scale
        b = s[i] / scale
        norm = np.sqrt(a ** 2 + b ** 2)
        c[i] = a / norm
        s[i] = b / norm
        alpha = scale * norm
    for j in range(p):
        xx = 0.0
        for ii in range(1, j + 2):
            i = j - ii + 1
            t = c[i] * xx + s[i] * r[i, j]
            r[i, j] = c[i] * r[i, j] - s[i] * xx
            xx = t
    return 0


# ------------------------------------------------------------------------------------------------ user models
class ExpReg:
    """testcases/mcmcrun.F90:48-122."""

    def __init__(self, x, y):
        self.x, self.y = np.asarray(x, float), np.asarray(y, float)

    def ssfunction(self, theta):
        return np.array([np.sum((self.y - theta[0] * np.exp(-theta[1] * self.x)) ** 2)])   # :89,104

    def checkbounds(self, theta):
        return not np.any(theta <= 0.0)                                                    # :112-122


class Gauss:
    """testcases/mcmcrun4.F90:47, checkbounds0.f90:3-18."""

    def __init__(self, mu, lam):
        self.mu, self.lam = np.asarray(mu, float), np.asarray(lam, float)

    def ssfunction(self, theta):
        return np.array([np.dot(self.lam @ (theta - self.mu), theta - self.mu)])

    def checkbounds(self, theta):
        return True


class Hier:
    """Build-defined hierarchical normal means (csrc/models.cuh HierN): not in the reference."""

    def __init__(self, y):
        self.y = np.asarray(y, float)   # groups x observations

    def ssfunction(self, theta):
        G = self.y.shape[0]
        mu, ltau = theta[G], theta[G + 1]
        itau2 = np.exp(-2.0 * ltau)
        tg = theta[:G]
        ss = np.sum((self.y - tg[:, None]) ** 2) + np.sum((tg - mu) ** 2 * itau2)
        return np.array([ss + 2.0 * G * ltau + mu * mu / 100.0 + ltau * ltau / 4.0])

    def checkbounds(self, theta):
        return True


# ------------------------------------------------------------------------------------------------ the sampler
NML_DEFAULTS = dict(  # mcmcinit.F90:184-230
    nsimu=0, doadapt=1, doburnin=0, burnintime=0, badaptint=-1, greedy=0, scalelimit=0.05, scalefactor=2.5,
    drscale=0.0, adaptint=100, adapthist=0, adaptend=0, initcmatn=0, N0=1.0, S02=0.0, updatesigma=1, condmax=0.0,
    method="dram", alphatarget=0.234, nuparam=0.7)


class Run:
    """module mcmcmod's globals (mcmc.F90:28-60) + the subroutines #included into it."""

    def __init__(self, nml, model, par0, cmat0, sigma2, nobs, uniforms, prior=None):
        self.__dict__.update(NML_DEFAULTS)
        self.__dict__.update(nml)
        self.method = str(self.method).lower()
        self.model = model
        self.prior = prior
        self.rng = Stream(uniforms)
        self.check_mcmcinit_parameters()
        self.status = 0
        # ---- MCMC_init.F90:45-154
        self.par0 = np.array(par0, dtype=np.float64)
        self.cmat0 = np.array(cmat0, dtype=np.float64)
        self.sigma2 = np.array(sigma2, dtype=np.float64).ravel()
        self.nobs = np.array(nobs).ravel()
        self.npar = self.par0.size
        self.nycol = self.sigma2.size
        n = self.npar
        self.R = np.zeros((n, n))
        self.R2 = np.zeros((n, n))
        self.iC = np.zeros((n, n))
        self.qcovstd = np.zeros(n)
        self.chaincmat = self.cmat0.copy()          # :99-102
        self.chainmean = self.par0.copy()
        self.chainwsum = float(self.initcmatn)
        info = self.MCMC_calculate_R(self.cmat0)    # :108
        if info != 0:
            raise RuntimeError("could not factor the initial covariance")
        if self.S02 <= 0.0:                          # :114-116
            self.S02 = float(self.sigma2[0])
        self.ncolchain = n + 1
        self.chain = np.zeros((self.nsimu, n + 1))
        self.sschain = np.zeros((self.nsimu, self.nycol + 1))
        self.s2chain = np.zeros((self.nsimu, self.nycol))
        self.chainind = 0                            # :147-154 (1-based row index, 0 = none yet)
        self.simuind = 1
        self.stayed = self.bndstayed = self.erstayed = self.draccepted = self.drtries = 0
        # MCMC_adapt's `save` variables, MCMC_adapt.F90:15,19
        self.lastind, self.lastfreq = 1, 0

    # mcmcinit.F90:235-368 (the parts that reach the sampler)
    def check_mcmcinit_parameters(self):
        if self.adapthist < 0:
            self.adapthist = 0
        if self.adaptint < 0:
            self.adaptint, self.doadapt = 0, 0
        if self.burnintime < 0:
            self.burnintime = 0
        if self.badaptint <= 0:
            self.badaptint = self.adaptint
        if self.badaptint == 0:
            self.doburnin = 0
        if self.initcmatn < 0:
            self.initcmatn = 0
        if self.scalelimit < 0.0 or self.scalelimit > 0.5:
            raise ValueError("scalelimit")
        if self.scalefactor < 0.0:
            self.scalefactor = 1.0
        if self.method == "scam":
            self.doscam = True
            if self.condmax <= 0.0:
                self.condmax = 1.0e15
            self.doburnin = 0
            self.drscale = 0.0
        else:
            self.doscam = False
        if self.method == "ram":
            self.drscale = 0.0
        self.dodr = self.drscale > 0.0
        self.usesvd = 1 if self.condmax > 0.0 else 0

    # ---- plugin wrappers, MCMC_DRAM.F90:37-90
    def MCMC_priorfun(self, par):
        if self.prior is None:
            return 0.0
        mu, sig = self.prior
        m = sig > 0.0                                # priorfun.f90:97-100
        return float(np.sum(((par[m] - mu[m]) / sig[m]) ** 2))

    # MCMC_DRAM.F90:20-31
    def MCMC_propose(self, oldpar, R):
        z = self.rng.random_normal(self.npar)
        if self.usesvd != 0:
            return oldpar + matmulx(R, z), z
        return oldpar + matmulu_t(R, z), z

    # MCMC_DRAM.F90:100-118
    def MCMC_alpha(self, ss1, sspri1, ss2, sspri2):
        tst = -0.5 * (np.sum((ss2 - ss1) / self.sigma2) + (sspri2 - sspri1))
        if tst >= 0.0:
            return 1.0
        if tst < LOG_REALMIN:
            return 0.0
        return float(np.exp(tst))

    # MCMC_DRAM.F90:140-155
    def MCMC_reject(self, alpha):
        reject = True
        if alpha >= 1.0:
            reject = False
        elif alpha > 0.0:
            u = self.rng.random_number(1)[0]
            if u <= alpha:
                reject = False
        return reject

    # MCMC_DRAM.F90:162-186
    def MCMC_DR_alpha13(self, oldpar, ss1, sspri1, newpar, ss2, sspri2, alpha12, newpar2, ss3, sspri3):
        with np.errstate(over="ignore", invalid="ignore"):
            if alpha12 == 0.0:
                alpha32 = 0.0
            else:
                tst32 = -0.5 * (np.sum((ss2 - ss3) / self.sigma2) + (sspri2 - sspri3))
                alpha32 = min(1.0, float(np.exp(tst32)))
            l2 = -0.5 * (np.sum((ss3 - ss1) / self.sigma2) + (sspri3 - sspri1))
            q1 = -0.5 * (np.sum(matmuls(self.iC, newpar2 - newpar) * (newpar2 - newpar))
                         - np.sum(matmuls(self.iC, oldpar - newpar) * (oldpar - newpar)))
            a = float(np.exp(l2 + q1)) * (1.0 - alpha32) / (1.0 - alpha12)
        if a != a:
            return a            # NaN: min(1,NaN) is processor dependent; NaN rejects without a draw either way
        return min(1.0, a)

    # MCMC_DRAM.F90:192-206
    def MCMC_updatesigma2(self, ss):
        if self.updatesigma != 0:
            for j in range(self.nycol):
                g = self.rng.random_gamma(self.N0 / 2.0 + float(self.nobs[j]) / 2.0, 2.0 / (self.N0 * self.S02 + ss[j]))
                self.sigma2[j] = 1.0 / g

    # MCMC_aux.F90:166-185 (save_method = 'memory'); chainind is 1-based like the Fortran
    def MCMC_savechain(self, par, ss, reject):
        n = self.npar
        if reject:
            self.chain[self.chainind - 1, n] += 1.0
        else:
            self.chainind += 1
            self.chain[self.chainind - 1, :n] = par
            self.chain[self.chainind - 1, n] = 1.0
            self.sschain[self.chainind - 1, :self.nycol] = ss
        self.sschain[self.chainind - 1, self.nycol] = self.chain[self.chainind - 1, n]
        if self.updatesigma != 0:
            self.s2chain[self.simuind - 1, :] = self.sigma2

    # MCMC_adapt.F90:181-230
    def MCMC_calculate_R(self, cmat):
        n = self.npar
        if self.doscam:
            R0, std, info = scam_svd(cmat, self.condmax)
            if R0 is None:
                return info
            if info == -1:
                info = 0
            self.qcovstd = std
            self.R = R0
            return info
        if self.usesvd != 0:
            R0, info = covtor_svd(cmat, self.condmax)
            if R0 is None:
                return info
            if info == -1:
                cmat[:, :] = R0 @ R0.T          # cmat is intent(inout): the caller's chaincmat changes
                info = 0
        else:
            R0, info = lapack.dpotrf(np.asfortranarray(cmat.copy()), lower=0, clean=0)   # covtor, matutils.F90:345-374
        if info != 0:
            return info
        self.R = R0 * 2.4 / np.sqrt(float(n))
        if self.dodr:
            iC, info2 = lapack.dpotri(np.asfortranarray(self.R.copy()), lower=0)
            if info2 != 0:
                raise RuntimeError("cannot invert cmat")
            self.iC = iC
            self.R2 = self.R / self.drscale
        return 0

    # MCMC_adapt.F90:12-174
    def MCMC_adapt(self, simuind):
        n = self.npar
        if self.doadapt == 0 and self.doburnin == 0:
            return
        if self.adaptend > 0 and simuind > self.adaptend:
            return
        ma = simuind % self.adaptint if self.adaptint != 0 else 1     # mod(i,0) is undefined; adaptint=0 => doadapt=0
        mb = simuind % self.badaptint if self.badaptint != 0 else 1
        if ma != 0 and mb != 0:
            return
        chain = self.chain
        if simuind < self.burnintime and self.doburnin != 0 and mb == 0:
            staypc = float(self.stayed) / float(simuind)
            if staypc > 1.0 - self.scalelimit:
                self.R = self.R / self.scalefactor
                if self.dodr:
                    self.R2 = self.R2 / self.scalefactor
                    self.iC = self.iC * self.scalefactor * self.scalefactor
                return
            elif staypc < self.scalelimit:
                self.R = self.R * self.scalefactor
                if self.dodr:
                    self.R2 = self.R2 * self.scalefactor
                    self.iC = self.iC / self.scalefactor / self.scalefactor
                return
            elif self.greedy != 0:
                self.chainwsum = float(self.initcmatn)
                self.chaincmat = self.cmat0.copy()
                self.chainmean = self.par0.copy()
                self.chaincmat, self.chainmean, self.chainwsum = covmat(
                    chain[0:self.chainind, 0:n], self.chaincmat, [1.0], self.chainmean, self.chainwsum, True)
                self.lastfreq = int(chain[self.chainind - 1, n])
            self.lastind = self.chainind
        elif simuind >= self.burnintime + self.adaptint + self.adapthist and self.doadapt != 0:
            if simuind == self.burnintime + self.adaptint + self.adapthist:
                self.chainwsum = float(self.initcmatn)
                self.chaincmat = self.cmat0.copy()
                self.chainmean = self.par0.copy()
            if self.adapthist > 1:
                istart = self.chainind
                histsum = int(chain[istart - 1, n])
                while histsum < self.adapthist and istart > 1:
                    istart -= 1
                    histsum += int(chain[istart - 1, n])
                newfreq = int(chain[istart - 1, n])
                chain[istart - 1, n] = float(newfreq - histsum + self.adapthist)
                self.chaincmat, self.chainmean, self.chainwsum = covmat(
                    chain[istart - 1:self.chainind, 0:n], self.chaincmat, chain[istart - 1:self.chainind, n].copy(),
                    self.chainmean, self.chainwsum, False)
                chain[istart - 1, n] = float(newfreq)
            else:
                newfreq = int(chain[self.lastind - 1, n])
                chain[self.lastind - 1, n] = float(newfreq - self.lastfreq)
                istart = self.lastind
                self.chaincmat, self.chainmean, self.chainwsum = covmat(
                    chain[istart - 1:self.chainind, 0:n], self.chaincmat, chain[istart - 1:self.chainind, n].copy(),
                    self.chainmean, self.chainwsum, True)
                chain[self.lastind - 1, n] = float(newfreq)
                self.lastfreq = int(chain[self.chainind - 1, n])
                self.lastind = self.chainind
        else:
            return
        info = self.MCMC_calculate_R(self.chaincmat)
        if info != 0:
            self.status |= 1      # "Warning: error in Chol/SVD, not adapting": old R kept

    # ---------------------------------------------------------------- MCMC_run.F90:12-114
    def _first_point(self):
        oldpar = self.par0.copy()
        sspri1 = self.MCMC_priorfun(oldpar)
        ss1 = self.model.ssfunction(oldpar)
        self.MCMC_savechain(oldpar, ss1, False)
        return oldpar, ss1, sspri1

    def MCMC_run(self):
        oldpar, ss1, sspri1 = self._first_point()
        for i in range(2, self.nsimu + 1):
            self.simuind = i
            newpar, _ = self.MCMC_propose(oldpar, self.R)
            inbounds = self.model.checkbounds(newpar)
            if not inbounds:
                if not self.dodr:
                    self.bndstayed += 1
                ss2 = np.full(self.nycol, HUGE)
                sspri2 = HUGE
                alpha12 = 0.0
                reject = True
            else:
                sspri2 = self.MCMC_priorfun(newpar)
                ss2 = self.model.ssfunction(newpar)
                alpha12 = self.MCMC_alpha(ss1, sspri1, ss2, sspri2)
                reject = self.MCMC_reject(alpha12)
            if reject and self.dodr:
                self.drtries += 1
                newpar2, _ = self.MCMC_propose(oldpar, self.R2)
                inbounds = self.model.checkbounds(newpar2)
                if not inbounds:
                    self.bndstayed += 1
                    reject = True
                else:
                    sspri3 = self.MCMC_priorfun(newpar2)
                    ss3 = self.model.ssfunction(newpar2)
                    alpha13 = self.MCMC_DR_alpha13(oldpar, ss1, sspri1, newpar, ss2, sspri2, alpha12, newpar2, ss3, sspri3)
                    reject = self.MCMC_reject(alpha13)
                    if not reject:
                        self.draccepted += 1
                        newpar, ss2, sspri2 = newpar2, ss3, sspri3
            if reject:
                self.stayed += 1
            else:
                ss1, sspri1, oldpar = ss2, sspri2, newpar
            self.MCMC_updatesigma2(ss1)
            self.MCMC_savechain(oldpar, ss1, reject)
            self.MCMC_adapt(i)
        self.oldpar, self.ss1 = oldpar, ss1

    # ---------------------------------------------------------------- MCMC_run_ram.F90:13-179
    def MCMC_run_ram(self):
        oldpar, ss1, sspri1 = self._first_point()
        alpha12 = 0.0     # undefined in the reference until the first in-bounds proposal: declared 0
        for i in range(2, self.nsimu + 1):
            self.simuind = i
            newpar, u = self.MCMC_propose(oldpar, self.R)       # MCMC_propose_ram keeps u, :87-101
            inbounds = self.model.checkbounds(newpar)
            if not inbounds:
                self.bndstayed += 1
                reject = True
            else:
                sspri2 = self.MCMC_priorfun(newpar)
                ss2 = self.model.ssfunction(newpar)
                alpha12 = self.MCMC_alpha(ss1, sspri1, ss2, sspri2)
                reject = self.MCMC_reject(alpha12)
            if reject:
                self.stayed += 1
            else:
                ss1, sspri1, oldpar = ss2, sspri2, newpar
            self.MCMC_updatesigma2(ss1)
            self.MCMC_savechain(oldpar, ss1, reject)
            self.MCMC_adapt_ram(i, u, alpha12)
        self.oldpar, self.ss1 = oldpar, ss1

    def MCMC_adapt_ram(self, simuind, u, alpha):       # :104-179
        if self.doadapt == 0:
            return
        if simuind < self.burnintime and self.doburnin != 0:
            return
        a = 1.0 / float(np.float32(simuind)) ** self.nuparam * (alpha - self.alphatarget)   # real(simuind): single
        if a >= 0.0:
            dchud(self.R, u / np.sum(u ** 2) * a)
        else:
            if dchdd(self.R, -u / np.sum(u ** 2) * a) != 0:
                self.status |= 2      # the reference stops here (matutils.F90:716-722); declared: flag and carry on

    # ---------------------------------------------------------------- MCMC_run_scam.F90:12-138
    def MCMC_run_scam(self):
        oldpar, ss1, sspri1 = self._first_point()
        n = self.npar
        for i in range(2, self.nsimu + 1):
            self.simuind = i
            rejall = True
            for j in range(n):
                rotpar = matmulx(self.R, oldpar, "t")                     # MCMC_scam_rotate 'f', :122-138
                z = self.rng.random_normal(1) * self.qcovstd[j]
                rotpar[j] = rotpar[j] + z[0]
                newpar = matmulx(self.R, rotpar, "n")
                inbounds = self.model.checkbounds(newpar)
                if not inbounds:
                    if not self.dodr:
                        self.bndstayed += 1
                    reject = True
                else:
                    sspri2 = self.MCMC_priorfun(newpar)
                    ss2 = self.model.ssfunction(newpar)
                    alpha12 = self.MCMC_alpha(ss1, sspri1, ss2, sspri2)
                    reject = self.MCMC_reject(alpha12)
                if not reject:
                    ss1, sspri1, oldpar = ss2, sspri2, newpar
                    rejall = False
            if rejall:
                self.stayed += 1
            self.sschain[self.chainind - 1, self.nycol] = self.chain[self.chainind - 1, n]
            self.MCMC_updatesigma2(ss1)
            self.MCMC_savechain(oldpar, ss1, rejall)
            self.MCMC_adapt(i)
        self.oldpar, self.ss1 = oldpar, ss1

    # ---------------------------------------------------------------- MCMC_run_er.F90:12-107
    def MCMC_run_er(self):
        self.dodr = False
        oldpar, ss1, sspri1 = self._first_point()
        for i in range(2, self.nsimu + 1):
            self.simuind = i
            newpar, _ = self.MCMC_propose(oldpar, self.R)
            inbounds = self.model.checkbounds(newpar)
            if not inbounds:
                self.bndstayed += 1
                reject = True
            else:
                u = self.rng.random_number(1)[0]                          # MCMC_sscrit, MCMC_DRAM.F90:124-135
                sscrit = -2.0 * np.log(u) + np.sum(ss1 / self.sigma2) + sspri1
                sspri2 = self.MCMC_priorfun(newpar)
                if sspri2 >= sscrit:
                    reject = True
                    self.erstayed += 1
                else:
                    sscrit = self.sigma2[0] * (sscrit - sspri2)
                    ss2 = self.model.ssfunction(newpar)                   # ssfunction_er0.f90 forwards to ssfunction
                    reject = bool(np.sum(ss2) >= sscrit)
            if reject:
                self.stayed += 1
            else:
                ss1, sspri1, oldpar = ss2, sspri2, newpar
            self.MCMC_updatesigma2(ss1)
            self.MCMC_savechain(oldpar, ss1, reject)
            self.MCMC_adapt(i)
        self.oldpar, self.ss1 = oldpar, ss1

    # mcmc_main.F90:29-37
    def run(self):
        {"scam": self.MCMC_run_scam, "er": self.MCMC_run_er, "ram": self.MCMC_run_ram}.get(self.method, self.MCMC_run)()
        return self

    def counters(self):
        return dict(stayed=self.stayed, bndstayed=self.bndstayed, draccepted=self.draccepted, drtries=self.drtries,
                    chainind=self.chainind, simuind=self.simuind, erstayed=self.erstayed, ndrawn=self.rng.n)

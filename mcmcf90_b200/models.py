"""Host-side helpers that pack user-model data into the blob the device models read
(layouts documented in csrc/models.cuh; the reference plugin loads the same data from
files on first call, e.g. data.dat at testcases/mcmcrun.F90:69-86)."""
import numpy as np


def blob_expreg(x, y):
    """[n, +-max|x| (negative when some x < 0), x[npad], y[npad]] for y = theta1*exp(-theta2*x) (testcases/mcmcrun.F90:104)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
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
    """[d, 0, mu[dpad], Lam[d*d]] for ss = (theta-mu)' Lam (theta-mu) (testcases/mcmcrun4.F90:47)."""
    mu = np.asarray(mu, dtype=np.float64).ravel()
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


def load_dat(path):
    """ASCII matrix reader of the reference (loaddata, matutils.F90:1007-1280): whitespace or
    comma separated numbers, lines starting with one of #%!Cc are comments."""
    rows = []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s[0] in "#%!Cc":
                continue
            rows.append([float(t) for t in s.replace(",", " ").split()])
    return np.array(rows, dtype=np.float64)

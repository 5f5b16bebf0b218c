"""Shared test inputs: the reference's shipped testcase (testcases/data.dat, mcmcpar.dat,
mcmccov.dat, mcmcsigma2.dat -- values restated here because /root/reference does not
exist on the GPU box) and seeded synthetic variants."""
import numpy as np

DATA_X = np.arange(11.0)
DATA_Y = np.array([9.33, 9.40, 8.99, 7.06, 7.13, 6.69, 4.69, 4.24, 4.77, 3.86, 4.02])
PAR0 = np.array([10.0, 0.1])
CMAT0 = np.diag([0.2, 0.001])
SIGMA2 = [0.5]
NOBS = [11]

# testcases/mcmcinit.nml as shipped (DR off, burn-in scaling only)
NML_SHIPPED = dict(nsimu=1000, doadapt=1, adaptint=200, burnintime=1000, doburnin=1, drscale=0.0,
                   updatesigma=1, N0=1.0, S02=0.0)
# DRAM variant of SURVEY.md 8d C1(ii)
NML_DRAM = dict(nsimu=2001, doadapt=1, adaptint=100, burnintime=0, doburnin=0, drscale=2.0, initcmatn=1,
                updatesigma=1, N0=1.0, S02=0.0)


def synth_expreg(n, seed=2024):
    """C3-style data: x on [0,10], y = 10 exp(-0.1 x) + N(0, 0.5)."""
    rng = np.random.default_rng(seed)
    x = 10.0 * np.arange(n) / (n - 1)
    y = 10.0 * np.exp(-0.1 * x) + rng.normal(0.0, np.sqrt(0.5), n)
    return x, y

"""Writes tests/golden/golden_c1.npz: outputs of the CPU restatement (oracle/) on the reference's shipped
testcase inputs (testcases/data.dat, mcmcpar.dat, mcmccov.dat, mcmcsigma2.dat -- values restated in
tests/cases.py) under injected uniform streams, for the four samplers.

The reference itself cannot run here (Fortran, no compiler) and ships no expected outputs, so these
vectors pin the ORACLE (against regressions and across machines/compilers) and let the GPU parity tests
compare against committed numbers; they do not pin the oracle to the reference ("parity unpinned",
DESIGN.md 2).  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases  # noqa: E402

NSIMU = 301
SEEDS = {"shipped": 1, "dram": 2, "ram": 3, "scam": 4, "scam_hier": 5, "er": 6, "ap": 7, "greedy": 8,
         "gauss_dram": 9, "gauss_ram": 10, "gauss_er": 11, "gauss_ap": 12, "gauss_greedy": 13}
HIER_Y = np.round(np.random.default_rng(99).normal(size=(4, 1)) + np.random.default_rng(98).normal(size=(4, 3)), 3)
CASES = {
    "shipped": dict(cases.NML_SHIPPED, nsimu=NSIMU, burnintime=150, adaptint=50),
    "dram": dict(cases.NML_DRAM, nsimu=NSIMU, adaptint=50),
    "ram": dict(method="ram", nsimu=NSIMU, updatesigma=1, N0=1.0, S02=0.0),
    "scam": dict(method="scam", nsimu=NSIMU, adaptint=50, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0),
    # SCAM on the hierarchical-means model (d = 6): the case the warp-per-chain GPU kernels are compared on
    "scam_hier": dict(method="scam", nsimu=NSIMU, adaptint=50, initcmatn=1, updatesigma=0),
    # early-rejection sampler (MCMC_run_er.F90), AP window (adapthist > 1) and greedy burn-in (MCMC_adapt.F90:83-136)
    "er": dict(method="er", nsimu=NSIMU, adaptint=100, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0),
    "ap": dict(cases.NML_DRAM, nsimu=NSIMU, adaptint=40, adapthist=60),
    "greedy": dict(nsimu=NSIMU, adaptint=50, burnintime=150, doburnin=1, badaptint=25, greedy=1, scalelimit=0.05,
                   drscale=2.0, initcmatn=1, updatesigma=1, N0=1.0, S02=0.0),
    # the same samplers on a 6-dim correlated Gaussian target (the quadratic form of testcases/mcmcrun4.F90:47): the
    # cases the warp-per-chain GPU kernels (run-time npar) are compared on
    "gauss_dram": dict(nsimu=NSIMU, adaptint=50, drscale=2.0, initcmatn=1, updatesigma=0),
    "gauss_ram": dict(method="ram", nsimu=NSIMU, updatesigma=0),
    "gauss_er": dict(method="er", nsimu=NSIMU, adaptint=100, initcmatn=3, updatesigma=1, N0=4.0, S02=1.0),
    "gauss_ap": dict(nsimu=NSIMU, adaptint=40, adapthist=60, drscale=2.0, initcmatn=1, updatesigma=0),
    "gauss_greedy": dict(nsimu=NSIMU, adaptint=50, burnintime=150, doburnin=1, badaptint=25, greedy=1, scalelimit=0.05,
                         drscale=2.0, initcmatn=18, updatesigma=0),
}
GAUSS_D = 6


def gauss_target(d, rho=0.9):
    sd = 1.0 + 9.0 * np.arange(d) / max(d - 1, 1)
    sig = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d))) * np.outer(sd, sd)
    lam = np.linalg.inv(sig)
    return np.zeros(d), 0.5 * (lam + lam.T)


def uniforms(name):
    """The injected stream of a case (regenerated from its seed, not stored)."""
    return np.random.default_rng(SEEDS[name]).random(40 * NSIMU)


def inputs(name):
    """(model_id, blob, par0, cmat0, sigma2, nobs) of a case."""
    if name.startswith("gauss_"):
        mu, lam = gauss_target(GAUSS_D)
        return O.MODEL_GAUSS, O.blob_gauss(mu, lam), np.zeros(GAUSS_D), 0.5 * np.eye(GAUSS_D), [1.0], [1]
    if name == "scam_hier":
        d = HIER_Y.shape[0] + 2
        return O.MODEL_HIER, O.blob_hier(HIER_Y), np.full(d, 0.1), 0.1 * np.eye(d), [1.0], [1]
    return O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), cases.PAR0, cases.CMAT0, cases.SIGMA2, cases.NOBS


def main():
    out = {}
    for name, nml in CASES.items():
        model_id, blob, par0, cmat0, sigma2, nobs = inputs(name)
        ch = O.Chain(O.make_cfg(**nml), model_id, blob, par0, cmat0, sigma2, nobs)
        ch.inject(uniforms(name))
        ch.run()
        r = ch.results()
        assert r["status"] == 0, (name, r["status"])
        for k in ("chain", "sschain", "s2chain", "R", "cmat", "mean", "par", "sigma2"):
            out["%s_%s" % (name, k)] = np.asarray(r[k])
        out[name + "_counters"] = np.array([r[k] for k in ("stayed", "bndstayed", "draccepted", "drtries", "chainind",
                                                          "simuind", "status", "ndrawn")], dtype=np.int64)
        out[name + "_wsum"] = np.array(r["wsum"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_c1.npz"), **out)
    print("wrote golden_c1.npz:", {k: v.shape for k, v in out.items() if k.endswith("_chain")})


if __name__ == "__main__":
    main()

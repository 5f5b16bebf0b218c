"""Run one golden case (tests/golden/make_golden.py) through the PATCHED REFERENCE EXECUTABLE built by
oracle/Makefile.ref:  python oracle/ref_case.py <case> <workdir> <path to mcmcrun_ref>

Writes the reference's own input files (mcmcinit.nml, data.dat, mcmcpar.dat, mcmccov.dat, mcmcsigma2.dat;
formats of initialize.F90:41-119) and uniforms.bin (the injected stream) into <workdir>, runs the executable there
and leaves chain.dat / sschain.dat / s2chain.dat for the caller.  Test infrastructure only."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases  # noqa: E402
from tests.golden import make_golden as G  # noqa: E402


def main(name, work, exe):
    nml = dict(G.CASES[name])
    lines = ["&mcmc"]
    for k, v in nml.items():
        lines.append(" %s = %s" % (k, "'%s'" % v if isinstance(v, str) else repr(v)))
    lines += [" verbosity = 0", " printint = 100000000", " chainfile = 'chain.dat'", " ssfile = 'sschain.dat'",
              " s2file = 's2chain.dat'", "/"]
    open(os.path.join(work, "mcmcinit.nml"), "w").write("\n".join(lines) + "\n")
    np.savetxt(os.path.join(work, "data.dat"), np.column_stack([cases.DATA_X, cases.DATA_Y]), fmt="%.17g")
    np.savetxt(os.path.join(work, "mcmcpar.dat"), cases.PAR0[None, :], fmt="%.17g")
    np.savetxt(os.path.join(work, "mcmccov.dat"), cases.CMAT0, fmt="%.17g")
    np.savetxt(os.path.join(work, "mcmcsigma2.dat"), np.array([cases.SIGMA2, cases.NOBS], dtype=float), fmt="%.17g")
    G.uniforms(name).astype("<f8").tofile(os.path.join(work, "uniforms.bin"))
    subprocess.check_call([os.path.abspath(exe)], cwd=work, stdout=subprocess.DEVNULL)


if __name__ == "__main__":
    main(*sys.argv[1:4])

import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcmcf90_b200 as mb
from oracle import oracle as O
from tests import cases
from tests.test_k1_parity import _gpu_run, _oracle_run, BLOB11
nml = dict(nsimu=500, adaptint=100, drscale=2.0, initcmatn=1, updatesigma=1)
cm0 = np.diag([4.0, 0.02])
N = 4
u = np.random.default_rng(5).random((N, 40 * 500))
if len(sys.argv) > 1:
    os.environ["ORC_TRACE"] = "1"
    r = _oracle_run(nml, 3, BLOB11, cases.PAR0, u=u, cmat0=cm0)
    sys.exit(0)
s = _gpu_run(nml, N, BLOB11, cases.PAR0, u=u, cmat0=cm0, steps=0, splits=[0])
rows = []
for i in range(2, 501):
    s.run(1)
    c = s.counters()
    rows.append((i, c["stayed"][3], c["bndstayed"][3], c["ndrawn"][3], s.fetch("par")[3], s.fetch("sigma2")[3, 0], c["drtries"][3], c["draccepted"][3]))
out = subprocess.run([sys.executable, __file__, "oracle"], capture_output=True, text=True).stderr
orc = {}
for line in out.splitlines():
    if line.startswith("step"):
        f = dict(t.split("=") for t in line.split()[1:] if "=" in t)
        orc[int(f["i"])] = line
for (i, st, bn, nd, th, s2, drt, dra) in rows:
    f = dict(t.split("=") for t in orc[i].split()[1:] if "=" in t)
    if int(f["stayed"]) != st or int(f["nd"]) != nd or int(f["bnd"]) != bn:
        print("FIRST DIVERGENCE at step", i)
        for j in range(i - 2, i + 2):
            print(" orc:", orc[j])
            g = rows[j - 2]
            print(" gpu: i=%d stayed=%d bnd=%d nd=%d th=%.17g %.17g s2=%.17g drt=%d dra=%d" % (g[0], g[1], g[2], g[3], g[4][0], g[4][1], g[5], g[6], g[7]))
        break
else:
    print("no divergence")

"""debug: which leg-2 configuration does the driver actually run?"""
import os, sys, shutil, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from tests import cases
from tests.test_host_driver import HOST, NML_SHIPPED, write_testcase
tmp = tempfile.mkdtemp()
leg1, leg2 = os.path.join(tmp, "leg1"), os.path.join(tmp, "leg2")
os.makedirs(leg1)
nml = NML_SHIPPED.replace("method = 'dram'", "method = 'dram'\n nmlffile = 'final.nml'\n covnfile = 'mcmccovn.dat'") \
                 .replace("drscale     = 0", "drscale     = 2.0").replace("burnintime  = 1000", "burnintime  = 0") \
                 .replace("doburnin    = 1", "doburnin    = 0").replace("nsimu       = 1000", "nsimu       = 300") \
                 .replace("adaptint    = 200", "adaptint    = 50\n initcmatn = 1")
write_testcase(leg1, nml, "&mcmcb nchains = 2, seed = 17, store_chains = 1 /\n")
print(subprocess.run([os.path.join(HOST, "mcmcb_main"), leg1], capture_output=True, text=True).stdout[-300:])
shutil.copytree(leg1, leg2)
for src, dst in (("mcmccovf.dat", "mcmccov.dat"), ("mcmcparf.dat", "mcmcpar.dat"), ("mcmcsigma2f.dat", "mcmcsigma2.dat"), ("final.nml", "mcmcinit.nml")):
    shutil.copy(os.path.join(leg2, src), os.path.join(leg2, dst))
print(open(os.path.join(leg2, "mcmcinit.nml")).read())
for f in ("mcmcpar.dat", "mcmccov.dat", "mcmcsigma2.dat", "mcmccovn.dat"):
    print(f, open(os.path.join(leg2, f)).read())
r = subprocess.run([os.path.join(HOST, "mcmcb_main"), leg2], capture_output=True, text=True)
print(r.stdout[-400:], r.stderr[-400:])
chain = np.loadtxt(os.path.join(leg2, "chain.dat"), ndmin=2)
print("driver rows", chain.shape, chain[:3])
par0 = np.loadtxt(os.path.join(leg2, "mcmcpar.dat")); cmat0 = np.loadtxt(os.path.join(leg2, "mcmccov.dat")); s2n = np.loadtxt(os.path.join(leg2, "mcmcsigma2.dat"))
n0 = int(np.loadtxt(os.path.join(leg2, "mcmccovn.dat")))
kw1 = dict(nsimu=300, doadapt=1, adaptint=50, burnintime=0, doburnin=0, drscale=2.0, updatesigma=1, N0=1.0)
for ic in (n0, n0 + 300, 1):
    for s02 in (0.5, 0.0):
        ch = O.Chain(O.make_cfg(initcmatn=ic, S02=s02, **kw1), O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), par0, cmat0, [s2n[0]], [int(s2n[1])])
        ch.philox(17, 0); ch.run(); r2 = ch.results()
        print("oracle initcmatn=%d S02=%g rows=%d" % (ic, s02, r2["chain"].shape[0]), r2["chain"][:3].tolist() if ic == n0 and s02 == 0.5 else "")

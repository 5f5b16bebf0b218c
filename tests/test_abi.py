"""CPU-side checks of the C-ABI library: it loads, exports every symbol that
include/mcmcb200.h declares, and its config logic mirrors mcmcinit.F90 (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import mcmcf90_b200 as mb
from mcmcf90_b200 import binding
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "mcmcb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mcmcb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = mb.load_library()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert sorted(binding.EXPORTS) == names


def test_config_struct_matches_header_size():
    # the ctypes mirror must have the C struct's size; create() rejects a wrong abi_version
    c = mb.default_config()
    assert c.abi_version == 3
    h = C.c_void_p()
    c.abi_version = 99
    assert mb.load_library().mcmcb_create(C.byref(c), C.byref(h)) == -1


def test_defaults_mirror_namelist_defaults():
    c = mb.default_config()
    o = O.make_cfg()
    for k in ("method", "nsimu", "doadapt", "adaptint", "adapthist", "adaptend", "initcmatn", "doburnin", "burnintime",
              "badaptint", "greedy", "scalelimit", "scalefactor", "drscale", "condmax", "N0", "S02", "updatesigma",
              "alphatarget", "nuparam"):
        assert getattr(c, k) == getattr(o, k), k
    assert (c.adaptint, c.scalelimit, c.scalefactor, c.alphatarget, c.nuparam) == (100, 0.05, 2.5, 0.234, 0.7)


@pytest.mark.parametrize("kw", [dict(method="scam", drscale=3.0, doburnin=1), dict(method="ram", drscale=2.0),
                                dict(adaptint=-5), dict(badaptint=0, adaptint=0, doburnin=1), dict(condmax=1e10),
                                dict(drscale=2.0, initcmatn=-3, burnintime=-1, adapthist=-2)])
def test_check_config_mirrors_reference_rules(kw):
    c = mb.default_config(**kw)
    o = O.make_cfg(**kw)
    dodr, doscam, usesvd = C.c_int(), C.c_int(), C.c_int()
    assert mb.load_library().mcmcb_check_config(C.byref(c), C.byref(dodr), C.byref(doscam), C.byref(usesvd)) == 0
    O.lib().orc_check_params(C.byref(o))
    assert (dodr.value, doscam.value, usesvd.value) == (o.dodr, o.doscam, o.usesvd)
    for k in ("doadapt", "adaptint", "adapthist", "initcmatn", "doburnin", "burnintime", "badaptint", "drscale",
              "condmax"):
        assert getattr(c, k) == getattr(o, k), k


def test_check_config_rejects_bad_scalelimit():
    c = mb.default_config(scalelimit=0.7)
    assert mb.load_library().mcmcb_check_config(C.byref(c), None, None, None) == -1


def test_check_config_early_rejection_and_bad_method():
    # "no dr with er" (MCMC_run_er.F90:24-27): the reference switches delayed rejection off when MCMC_run_er starts;
    # here check_config does it.  An unknown method code is an error (the reference falls back to MCMC_run for unknown
    # method *strings*, which the namelist reader of host/ reproduces before this call).
    c = mb.default_config(method="er", drscale=2.0)
    dodr = C.c_int(7)
    assert mb.load_library().mcmcb_check_config(C.byref(c), C.byref(dodr), None, None) == 0
    assert (c.method, c.drscale, dodr.value) == (mb.ER, 0.0, 0)
    c = mb.default_config()
    c.method = 9
    assert mb.load_library().mcmcb_check_config(C.byref(c), None, None, None) == -1

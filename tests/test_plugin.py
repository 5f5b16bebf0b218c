"""User-model plugins (include/mcmcb200_plugin.cuh): a model compiled OUTSIDE the library registers itself by
name and runs through the same kernels -- the run-time form of the reference's link-time override of
ssfunction / priorfun / checkbounds (external_inc.h:4-28).  tests/plugin/user_models.cu is built by
__graft_entry__.build()."""
import os

import numpy as np
import pytest

import mcmcf90_b200 as mb
from oracle import oracle as O
from tests import cases

PLUGIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "plugin", "libuser_models.so")


def test_missing_plugin_is_an_error():
    with pytest.raises(mb.MCMCBError):
        mb.load_plugin("/nonexistent/libnomodel.so")


def test_plugin_registers_its_models_by_name():
    # no GPU needed: loading runs the plugin's registrars; create() then gets past the model lookup
    # (ENOMODEL for an unknown name) and only fails at the CUDA device on a CPU-only box
    assert os.path.exists(PLUGIN)
    mb.load_plugin(PLUGIN)
    from tests.conftest import has_gpu
    for name in ("user_expreg", "user_isogauss"):
        try:
            mb.Sampler(mb.default_config(nsimu=10, nchains=2, model=name)).close()
        except mb.MCMCBError as e:
            assert not has_gpu() and "ECUDA" in str(e), e
    with pytest.raises(mb.MCMCBError, match="ENOMODEL"):
        mb.Sampler(mb.default_config(nsimu=10, nchains=2, model="not_registered"))


@pytest.mark.gpu
def test_plugin_register_kernel_model_matches_oracle():
    mb.load_plugin(PLUGIN)
    nml = dict(cases.NML_DRAM, nsimu=301, adaptint=50)
    N = 48
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(1).normal(size=(N, 2)))
    blob = np.concatenate([[len(cases.DATA_X)], cases.DATA_X, cases.DATA_Y])  # the plugin's own layout
    s = mb.Sampler(mb.default_config(nchains=N, seed=21, model="user_expreg", **nml))
    s.set_data(blob)
    s.set_initial(par0, cases.CMAT0, cases.SIGMA2, cases.NOBS)
    s.run(300)
    ref = O.run_batch(O.make_cfg(**nml), O.MODEL_EXPREG, O.blob_expreg(cases.DATA_X, cases.DATA_Y), par0, cases.CMAT0,
                      cases.SIGMA2, cases.NOBS, seed=21, chain0=0, nthreads=4)
    cnt = s.counters()
    assert np.array_equal(cnt["stayed"], ref["counters"][:, 0]) and np.array_equal(cnt["drtries"], ref["counters"][:, 3])
    np.testing.assert_allclose(s.fetch("par"), ref["par"], rtol=1e-10)
    assert s.info()["kernel"] == 1
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["dram", "ram", "scam"])
def test_plugin_warp_kernel_model_matches_oracle(method):
    mb.load_plugin(PLUGIN)
    d, N = 7, 24
    nml = dict(method=method, nsimu=201, adaptint=50, initcmatn=1, updatesigma=0, drscale=2.0 if method == "dram" else 0.0)
    par0 = 0.1 * np.random.default_rng(3).normal(size=(N, d))
    s = mb.Sampler(mb.default_config(nchains=N, seed=5, model="user_isogauss", **nml))
    s.set_data(np.zeros(2))
    s.set_initial(par0, 0.2 * np.eye(d), [1.0], [1])
    s.run(200)
    ref = O.run_batch(O.make_cfg(**nml), O.MODEL_GAUSS, O.blob_gauss(np.zeros(d), np.eye(d)), par0, 0.2 * np.eye(d), [1.0],
                      [1], seed=5, chain0=0, nthreads=4)
    cnt = s.counters()
    same = cnt["stayed"] == ref["counters"][:, 0]
    assert same.mean() >= (0.9 if method == "ram" else 1.0)  # RAM: dchdd amplifies the lane-order rounding of ss
    tol = 1e-7 if method == "ram" else 1e-9
    np.testing.assert_allclose(s.fetch("par")[same], ref["par"][same], rtol=tol, atol=tol)
    assert s.info()["kernel"] == 2 and (cnt["status"] == 0).all()
    s.close()

"""Accuracy of the device FP64 exp used by model code (include/mcmcb200_model.cuh) against
mpmath (correctly rounded reference).  The bar for model arithmetic is north_star's 1e-12
relative on ss; the fast exps are held to 1.02 ulp (mcmcb_exp_fast: half an ulp each from the
correctly rounded table entry and the final DFMA, 0.01 from the polynomial) and to
1 ulp + |x s| 2^-52 (mcmcb_expmul_fast, whose only extra error is the rounding of the
pre-scaled factor)."""
import mpmath as mp
import numpy as np
import pytest

import mcmcf90_b200 as mb

pytestmark = pytest.mark.gpu


def ulp_err(got, a):
    mp.mp.dps = 40
    errs = []
    for g, x in zip(got, a):
        ref = mp.exp(mp.mpf(float(x)))
        u = np.spacing(float(ref))
        errs.append(abs(float((mp.mpf(float(g)) - ref) / mp.mpf(float(u)))))
    return np.array(errs)


def test_exp_fast_about_one_ulp():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-700, 700, 3000), rng.uniform(-1, 1, 3000), rng.uniform(-1e-3, 1e-3, 500),
                        [0.0, -0.0, 1.0, -1.0, 707.9, -707.9, 1e-300, -1e-300]])
    f, _ = mb.exp_selftest(a, 1.0)
    e = ulp_err(f, a)
    assert e.max() <= 1.02, e.max()


@pytest.mark.parametrize("scale", [-0.1, -0.1003, 3.7, -42.0])
def test_expmul_fast_error_bound(scale):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 10, 4000), [0.0, 10.0, 1e-9]])
    _, m = mb.exp_selftest(x, scale)
    mp.mp.dps = 40
    worst = 0.0
    for g, xi in zip(m, x):
        ref = mp.exp(mp.mpf(float(xi)) * mp.mpf(float(scale)))
        rel = abs(float((mp.mpf(float(g)) - ref) / ref))
        bound = 2.0 ** -52 * (1.0 + abs(float(xi) * scale))
        worst = max(worst, rel / bound)
    assert worst <= 1.0, worst


def test_out_of_range_falls_back_to_libm():
    a = np.array([-800.0, 800.0, -745.0, 709.0, np.nan, np.inf, -np.inf])
    f, m = mb.exp_selftest(a, 1.0)
    with np.errstate(over="ignore"):
        ref = np.exp(a)
    for got in (f, m):
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        np.testing.assert_allclose(got[ok], ref[ok], rtol=2e-16, atol=5e-324)

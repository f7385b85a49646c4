# -*- coding: utf-8 -*-
"""CUDA path vs the committed golden vectors produced by the reference package itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def test_kernel_values_and_gradients(gpu):
    from test_oracle_kernels import _golden_kernels
    for name, k in _golden_kernels().items():
        x1, x2 = GOLD[name + "__x1"], GOLD[name + "__x2"]
        rtol = 5e-12 if name == "ratquad" else 1e-13
        np.testing.assert_allclose(k.get_value(x1, x2), GOLD[name + "__value"], rtol=rtol, atol=1e-15)
        np.testing.assert_allclose(k.get_value(x1), GOLD[name + "__sym"], rtol=rtol, atol=1e-15)
        np.testing.assert_allclose(k.get_gradient(x1, x2, include_frozen=True), GOLD[name + "__grad"], rtol=1e-11,
                                   atol=1e-14)


def test_basic_solver_golden(gpu):
    import george_b200 as george
    from george_b200 import kernels
    x = GOLD["solver300__x"]
    s = george.BasicSolver(1.0 * kernels.ExpSquaredKernel(1.0))
    s.compute(x[:, None], np.ones(len(x)))
    y = np.sin(x)
    assert abs(s.log_determinant - float(GOLD["solver300__logdet"])) <= 1e-10 * abs(float(GOLD["solver300__logdet"]))
    np.testing.assert_allclose(s.apply_inverse(y), GOLD["solver300__alpha"], rtol=1e-9, atol=1e-12)
    assert abs(s.dot_solve(y) - float(GOLD["solver300__dot"])) <= 1e-10 * abs(float(GOLD["solver300__dot"]))


def test_predict_golden(gpu):
    import george_b200 as george
    from george_b200 import kernels
    for solver, kw in ((george.BasicSolver, {}), (george.HODLRSolver, {"tol": 1e-12})):
        kern = kernels.ExpSquaredKernel(1.0)
        kern.freeze_all_parameters()
        gp = george.GP(kern, white_noise=0.0, solver=solver, **kw)
        xs = GOLD["predict__x"]
        gp.compute(xs)
        mu, cov = gp.predict(np.sin(xs), GOLD["predict__x0"])
        # north star: predictive mean within 1e-6 relative
        assert np.max(np.abs(mu - GOLD["predict__mu"])) <= 1e-6 * np.max(np.abs(GOLD["predict__mu"]))
        assert np.allclose(cov, GOLD["predict__cov"], atol=1e-8)


def test_config4_miniature_and_docs_value(gpu):
    import george_b200 as george
    from george_b200 import kernels
    gp = george.GP(1.0 * kernels.Matern52Kernel(0.5, ndim=3))
    gp.compute(GOLD["cfg4mini__x"], 0.1)
    ll = gp.log_likelihood(GOLD["cfg4mini__y"])
    assert abs(ll - float(GOLD["cfg4mini__loglike"])) <= 1e-9 * abs(float(GOLD["cfg4mini__loglike"]))
    gph = george.GP(1.0 * kernels.Matern52Kernel(0.5, ndim=3), solver=george.HODLRSolver, tol=1e-12)
    gph.compute(GOLD["cfg4mini__x"], 0.1)
    assert abs(gph.log_likelihood(GOLD["cfg4mini__y"]) - float(GOLD["cfg4mini__loglike"])) <= 1e-6 * abs(float(GOLD["cfg4mini__loglike"]))

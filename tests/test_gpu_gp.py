# -*- coding: utf-8 -*-
"""GP-level parity, modelled on the reference's tests/test_gp.py, test_tutorial.py, test_pickle.py."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solvers():
    import george_b200 as george
    return [(george.BasicSolver, {}), (george.HODLRSolver, {"tol": 1e-10})]


@pytest.mark.parametrize("which", [0, 1])
def test_prediction(gpu, which):
    """reference tests/test_gp.py:59-83"""
    import george_b200 as george
    from george_b200 import kernels
    solver, kw = _solvers()[which]
    np.random.seed(42)
    kernel = kernels.ExpSquaredKernel(1.0)
    kernel.freeze_all_parameters()
    gp = george.GP(kernel, solver=solver, white_noise=0.0, **kw)
    x0 = np.linspace(-10, 10, 500)
    x = np.sort(np.random.uniform(-10, 10, 300))
    gp.compute(x)
    y = np.sin(x)
    mu0, cov0 = gp.predict(y, x0)
    Kstar = np.exp(-0.5 * (x0[:, None] - x[None, :]) ** 2)
    K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2)
    K[np.diag_indices_from(K)] += 1.0
    mu = np.dot(Kstar, np.linalg.solve(K, y))
    assert np.allclose(mu, mu0)
    mu1, var = gp.predict(y, x0, return_var=True)
    assert np.allclose(mu1, mu0)
    assert np.allclose(var, np.diag(cov0))
    cov = np.exp(-0.5 * (x0[:, None] - x0[None, :]) ** 2) - Kstar @ np.linalg.solve(K, Kstar.T)
    assert np.allclose(cov, cov0)


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("white_noise", [None, 0.1])
def test_gradient(gpu, which, white_noise):
    """reference tests/test_gp.py:16-56: grad_log_likelihood vs centred finite differences."""
    import george_b200 as george
    from george_b200 import kernels
    solver, kw = _solvers()[which]
    if solver is george.HODLRSolver:
        kw = {"tol": 1e-8}
    np.random.seed(123)
    N, ndim, eps = 305, 3, 1.32e-3
    kernel = 0.1 * kernels.ExpSquaredKernel(0.5, ndim=ndim)
    kwargs = dict(kw)
    if white_noise is not None:
        kwargs = dict(white_noise=white_noise, fit_white_noise=True, **kw)
    gp = george.GP(kernel, solver=solver, **kwargs)
    x = np.random.rand(N, ndim)
    x = x[np.argsort(x[:, 0])]
    y = np.sin(np.sum(x, axis=1))
    gp.compute(x, yerr=0.1)
    g0 = gp.grad_log_likelihood(y)
    vector = gp.get_parameter_vector()
    for i, v in enumerate(vector):
        vector[i] = v + eps
        gp.set_parameter_vector(vector)
        lp = gp.log_likelihood(y)
        vector[i] = v - eps
        gp.set_parameter_vector(vector)
        lm = gp.log_likelihood(y)
        vector[i] = v
        gp.set_parameter_vector(vector)
        grad = 0.5 * (lp - lm) / eps
        assert np.abs(grad - g0[i]) < 5 * eps, "grad {0}: {1} vs {2}".format(i, grad, g0[i])


@pytest.mark.parametrize("which", [0, 1])
def test_apply_inverse(gpu, which):
    """reference tests/test_gp.py:123-149"""
    import george_b200 as george
    from george_b200 import kernels
    solver, kw = _solvers()[which]
    np.random.seed(1234)
    x = np.sort(np.random.rand(201))
    y = np.sin(x)
    kernel = 0.1 * kernels.ExpSquaredKernel(0.5)
    gp = george.GP(kernel, solver=solver, **kw)
    gp.compute(x, yerr=0.1)
    K = gp.get_matrix(x)
    K[np.diag_indices_from(K)] += 0.1 ** 2 + 1.25e-12
    b = gp.apply_inverse(y)
    assert np.allclose(b, np.linalg.solve(K, y))
    y2 = np.vstack([y] * 5).T
    b2 = gp.apply_inverse(y2)
    assert np.allclose(b2, np.linalg.solve(K, y2))


def test_tutorial_basic_equals_hodlr(gpu):
    """reference tests/test_tutorial.py:12-43: N=50 < 2*min_size => a single leaf => HODLR is exact."""
    import george_b200 as george
    from george_b200 import kernels

    def model(params, t):
        _, _, amp, loc, sig2 = params
        return amp * np.exp(-0.5 * (t - loc) ** 2 / sig2)

    def lnlike(p, t, y, yerr, solver=george.BasicSolver):
        a, tau = np.exp(p[:2])
        gp = george.GP(a * kernels.Matern32Kernel(tau) + 0.001, solver=solver)
        gp.compute(t, yerr)
        return gp.log_likelihood(y - model(p, t))

    np.random.seed(1234)
    x = np.sort(np.random.rand(50))
    yerr = 0.05 + 0.01 * np.random.rand(len(x))
    y = np.sin(x) + yerr * np.random.randn(len(x))
    p = [0, 0, -1.0, 0.1, 0.4]
    l1 = lnlike(p, x, y, yerr)
    l2 = lnlike(p, x, y, yerr, solver=george.HODLRSolver)
    assert np.isfinite(l1)
    assert np.allclose(l1, l2)
    assert abs(l1 - l2) <= 1e-9 * abs(l1)


@pytest.mark.parametrize("which", [0, 1])
def test_pickle(gpu, which):
    """reference tests/test_pickle.py:21-36: a pickled GP must give the same answers (it refactorises lazily)."""
    import george_b200 as george
    from george_b200 import kernels
    solver, kw = _solvers()[which]
    np.random.seed(5)
    x = np.sort(np.random.rand(250))
    y = np.sin(x)
    gp = george.GP(0.5 * kernels.Matern32Kernel(0.3), solver=solver, **kw)
    gp.compute(x, 0.05)
    ll = gp.log_likelihood(y)
    gp2 = pickle.loads(pickle.dumps(gp, -1))
    assert np.allclose(gp2.log_likelihood(y), ll)
    k2 = pickle.loads(pickle.dumps(gp.kernel.kernel))
    assert np.allclose(k2.value_symmetric(x[:, None]), gp.get_matrix(x))


def test_parameter_change_triggers_recompute(gpu):
    import george_b200 as george
    from george_b200 import kernels
    np.random.seed(6)
    x = np.sort(np.random.rand(120))
    y = np.sin(5 * x)
    gp = george.GP(1.0 * kernels.ExpSquaredKernel(0.1), solver=george.HODLRSolver, tol=1e-12)
    gp.compute(x, 0.1)
    l0 = gp.log_likelihood(y)
    p = gp.get_parameter_vector()
    gp.set_parameter_vector(p + 0.3)
    assert not gp.computed
    l1 = gp.log_likelihood(y)
    assert gp.computed and l1 != l0
    gp.set_parameter_vector(p)
    assert np.allclose(gp.log_likelihood(y), l0)

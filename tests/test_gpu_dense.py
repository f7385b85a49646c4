# -*- coding: utf-8 -*-
"""BasicSolver parity (reference tests/test_solvers.py:29-58 + properties at larger N).  Tolerances are the
reference's own (np.allclose: rtol 1e-5, atol 1e-8) plus tighter checks against LAPACK on the oracle-built matrix."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_basic_solver_reference_case(gpu, oracle):
    import george_b200 as george
    from george_b200 import kernels
    from george_b200._spec import flatten
    np.random.seed(1234)
    N = 300
    x = np.sort(10 * np.random.randn(N))[:, None]
    yerr = np.ones(N)
    kernel = 1.0 * kernels.ExpSquaredKernel(1.0)
    solver = george.BasicSolver(kernel)
    solver.compute(x, yerr)
    K = kernel.get_value(x)
    K[np.diag_indices_from(K)] += yerr ** 2
    sgn, lndet = np.linalg.slogdet(K)
    assert sgn == 1.0
    assert np.allclose(solver.log_determinant, lndet)
    assert abs(solver.log_determinant - lndet) < 1e-9 * abs(lndet)
    y = np.sin(x[:, 0])
    b = solver.apply_inverse(y)
    assert b.shape == (N,)
    assert np.allclose(b, np.linalg.solve(K, y))
    assert np.allclose(solver.apply_inverse(K), np.eye(N))
    assert np.allclose(solver.get_inverse(), np.linalg.inv(K))
    assert np.allclose(solver.dot_solve(y), y @ np.linalg.solve(K, y))
    # apply_sqrt: r @ U with U^T U = K
    r = np.random.randn(4, N)
    U = np.linalg.cholesky(K).T
    assert np.allclose(solver.apply_sqrt(r), r @ U)
    assert np.allclose(solver.apply_sqrt(r[0]), r[0] @ U)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 129, 1000])
def test_sizes(gpu, oracle, n):
    import george_b200 as george
    from george_b200 import kernels
    from george_b200._spec import flatten
    rng = np.random.default_rng(n)
    x = rng.uniform(0, 1, (n, 3))
    yerr = 0.1 * np.ones(n)
    kernel = 1.0 * kernels.Matern52Kernel(0.5, ndim=3)
    K = oracle.value_symmetric(flatten(kernel), x) + np.diag(yerr ** 2)
    s = george.BasicSolver(kernel)
    s.compute(x, yerr)
    ld = np.linalg.slogdet(K)[1]
    assert abs(s.log_determinant - ld) <= 1e-10 * max(1.0, abs(ld))
    y = rng.normal(size=n)
    assert np.allclose(s.apply_inverse(y), np.linalg.solve(K, y), rtol=1e-8, atol=1e-10)
    Y = rng.normal(size=(n, 5))
    assert np.allclose(s.apply_inverse(Y), np.linalg.solve(K, Y), rtol=1e-8, atol=1e-10)


def test_not_positive_definite_raises_linalgerror(gpu):
    """scipy raises LinAlgError on a non-PD matrix; GP.recompute(quiet=True) depends on that type (gp.py:352-359)."""
    import george_b200 as george
    from george_b200 import kernels
    x = np.linspace(0, 1, 50)[:, None]
    kernel = kernels.CosineKernel(log_period=0.0) * kernels.DotProductKernel() + (-5.0 + 0 * 1) * 0 + kernels.DotProductKernel()
    kernel = kernels.DotProductKernel()  # rank-1 matrix: not PD once yerr = 0
    s = george.BasicSolver(kernel)
    with pytest.raises(np.linalg.LinAlgError):
        s.compute(x, np.zeros(50))
    gp = george.GP(kernel, white_noise=-80.0)
    gp.compute(x[:2], 1.0)
    gp._x = x
    gp._yerr2 = np.zeros(50)
    gp.computed = False
    assert gp.log_likelihood(np.ones(50), quiet=True) == -np.inf


def test_large_properties(gpu):
    """N = 8192 3-D Matern-5/2 (a quarter of config 4 per side): K (K^-1 y) == y and log-det vs LAPACK."""
    import george_b200 as george
    from george_b200 import kernels
    rng = np.random.default_rng(1)
    n = 8192
    x = rng.uniform(0, 1, (n, 3))
    x = x[np.argsort(x[:, 0])]
    yerr = 0.1 * np.ones(n)
    kernel = 1.0 * kernels.Matern52Kernel(0.5, ndim=3)
    s = george.BasicSolver(kernel)
    s.compute(x, yerr)
    K = kernel.get_value(x)
    K[np.diag_indices_from(K)] += yerr ** 2
    y = rng.normal(size=n)
    a = s.apply_inverse(y)
    assert np.linalg.norm(K @ a - y) <= 1e-9 * np.linalg.norm(y)
    import scipy.linalg
    c = scipy.linalg.cholesky(K, lower=True)
    ld = 2 * np.sum(np.log(np.diag(c)))
    assert abs(s.log_determinant - ld) <= 1e-10 * abs(ld)

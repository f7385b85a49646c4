# -*- coding: utf-8 -*-
"""K1/K2 parity: CUDA kernel-matrix build vs the CPU oracle (which is itself pinned to the reference's compiled
kernel_interface in test_oracle_kernels.py).  Tolerance: CUDA's exp/sin/cos/pow are within 2 ulp of glibc's, and
nvcc contracts a*b+c into FMA, so values agree to a few ulp: rtol 1e-13 (+ atol 1e-15 for values near zero)."""
import numpy as np
import pytest

from conftest import make_kernels

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-13, 1e-15


@pytest.mark.parametrize("name,kernel", make_kernels())
def test_values_match_oracle(gpu, oracle, name, kernel):
    from george_b200._spec import flatten
    rng = np.random.default_rng(7)
    nd = kernel.ndim
    x1 = rng.normal(size=(203, nd))
    x2 = rng.normal(size=(131, nd))
    spec = flatten(kernel)
    # pow() is 2 ulp on the device and the three-term sum can cancel: allow a few more ulp for the pow-based kernels
    rtol = 5e-12 if ("poly" in name or "ratquad" in name) else RTOL
    np.testing.assert_allclose(kernel.get_value(x1, x2), oracle.value_general(spec, x1, x2), rtol=rtol, atol=ATOL)
    ks = kernel.get_value(x1)
    np.testing.assert_allclose(ks, oracle.value_symmetric(spec, x1), rtol=rtol, atol=ATOL)
    assert np.array_equal(ks, ks.T)
    np.testing.assert_allclose(kernel.get_value(x1[:131], x2, diag=True), oracle.value_diagonal(spec, x1[:131], x2),
                               rtol=rtol, atol=ATOL)


@pytest.mark.parametrize("name,kernel", make_kernels())
def test_gradients_match_oracle(gpu, oracle, name, kernel):
    from george_b200._spec import flatten
    if len(kernel) == 0:
        pytest.skip("no parameters")
    rng = np.random.default_rng(11)
    nd = kernel.ndim
    x1 = rng.normal(size=(37, nd))
    x2 = rng.normal(size=(29, nd))
    spec = flatten(kernel)
    which = np.ones(kernel.full_size, dtype=np.uint32)
    g = kernel.get_gradient(x1, x2, include_frozen=True)
    np.testing.assert_allclose(g, oracle.gradient_general(spec, which, x1, x2), rtol=1e-12, atol=1e-14)
    gs = kernel.get_gradient(x1, include_frozen=True)
    np.testing.assert_allclose(gs, oracle.gradient_general(spec, which, x1, x1), rtol=1e-12, atol=1e-14)
    # frozen parameters are dropped from the last axis (kernels.py:115-127)
    name0 = kernel.get_parameter_names()[0]
    kernel.freeze_parameter(name0)
    assert kernel.get_gradient(x1).shape == (37, 37, kernel.full_size - 1)
    kernel.thaw_parameter(name0)


def test_edge_shapes(gpu, oracle):
    from george_b200 import kernels as K
    from george_b200._spec import flatten
    k = 1.0 * K.ExpSquaredKernel(1.0)
    spec = flatten(k)
    rng = np.random.default_rng(0)
    for n1, n2 in [(1, 1), (1, 300), (65, 1), (64, 128), (63, 129), (257, 3)]:
        x1, x2 = rng.normal(size=(n1, 1)), rng.normal(size=(n2, 1))
        np.testing.assert_allclose(k.get_value(x1, x2), oracle.value_general(spec, x1, x2), rtol=RTOL, atol=ATOL)
    for n in (1, 2, 63, 64, 65, 200):
        x = rng.normal(size=(n, 1))
        np.testing.assert_allclose(k.get_value(x), oracle.value_symmetric(spec, x), rtol=RTOL, atol=ATOL)
    with pytest.raises(RuntimeError):
        k.get_value(rng.normal(size=(5, 2)))


def test_x_gradient_rejects_too_many_dimensions(gpu):
    """kernel_x_gradient keeps ndim-vectors in BGP_MAX_DIM-sized local arrays: a wider input must be rejected with a
    ValueError, not written past them."""
    from george_b200 import kernels as K
    k = K.ExpSine2Kernel(gamma=1.0, log_period=0.3, ndim=10, axes=[0, 1])
    x = np.random.default_rng(0).normal(size=(7, 10))
    assert np.all(np.isfinite(k.get_value(x)))          # values are fine at any ndim
    with pytest.raises(ValueError):
        k.get_x1_gradient(x, x)
    with pytest.raises(ValueError):
        k.get_x2_gradient(x, x)


def test_user_kernels_from_yaml(gpu):
    """The two kernels generated from kernels/*.yml (tools/generate_kernels.py) evaluated on the device: values against
    their closed forms, hyper-parameter and input gradients against centred finite differences, mixed with built-in
    kernels in sums / products, and used by both solvers."""
    import george_b200 as george
    from george_b200 import kernels as K
    rng = np.random.default_rng(11)
    x = np.sort(rng.uniform(0, 8, 60))[:, None]
    k = K.CauchyKernel(metric=2.0)
    d2 = (x - x.T) ** 2
    np.testing.assert_allclose(k.get_value(x), 1.0 / (1.0 + d2 / 2.0), rtol=1e-13)
    P, L2 = 1.7, 3.0
    kd = K.DampedCosineKernel(log_period=np.log(P), log_decay=np.log(L2))
    d = x - x.T
    np.testing.assert_allclose(kd.get_value(x), np.exp(-d * d / (2 * L2)) * np.cos(2 * np.pi * d / P), rtol=1e-12, atol=1e-14)
    x2 = rng.uniform(0, 3, (25, 2))
    for kern in (K.CauchyKernel(metric=[1.0, 0.5], ndim=2), 0.7 * K.CauchyKernel(1.3, ndim=2) + kd_nd(K),
                 K.DampedCosineKernel(log_period=0.2, log_decay=0.4, ndim=2, axes=1) * K.ExpSquaredKernel(2.0, ndim=2)):
        kern.test_gradient(x2)
        kern.test_x1_gradient(x2)
        kern.test_x2_gradient(x2)
    # the generated kernels behind the solvers: log-likelihood vs dense numpy
    n = 1500
    xs = np.sort(rng.uniform(0, 15, n))
    yerr = 0.1 * np.ones(n)
    y = np.sin(xs) + 0.1 * rng.normal(size=n)
    kern = 1.0 * K.CauchyKernel(metric=1.0) + 0.5 * K.DampedCosineKernel(log_period=np.log(3.0), log_decay=np.log(20.0))
    Kd = kern.get_value(xs[:, None]) + np.diag(yerr ** 2 + george.gp.TINY)
    ll_ref = -0.5 * (n * np.log(2 * np.pi) + np.linalg.slogdet(Kd)[1]) - 0.5 * y @ np.linalg.solve(Kd, y)
    for solver, kw in ((george.BasicSolver, {}), (george.HODLRSolver, dict(tol=1e-12, min_size=100))):
        gp = george.GP(kern, solver=solver, **kw)
        gp.compute(xs, yerr)
        assert abs(gp.log_likelihood(y) - ll_ref) <= 1e-8 * abs(ll_ref)


def kd_nd(K):
    return K.DampedCosineKernel(log_period=0.1, log_decay=0.7, ndim=2, axes=[0, 1])

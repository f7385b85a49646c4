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

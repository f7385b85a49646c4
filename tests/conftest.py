# -*- coding: utf-8 -*-
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        from george_b200 import _lib
        return _lib.load().bgp_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle
    _oracle.lib()
    return _oracle


@pytest.fixture(scope="session")
def gpu():
    assert _have_gpu(), "no sm_100 device / libbgp_b200.so missing: GPU tests cannot run (no CPU fallback exists)"
    return True


def make_kernels():
    """Kernel zoo covering every kernel type, metric type, axes subsets, blocks, sums and products."""
    from george_b200 import kernels as K
    zoo = [
        ("expsq_1d", 1.0 * K.ExpSquaredKernel(1.0)),
        ("m32_1d", 2.3 * K.Matern32Kernel(0.7)),
        ("m52_3d_iso", K.Matern52Kernel(0.5, ndim=3)),
        ("m52_3d_axis", K.Matern52Kernel([0.5, 1.0, 2.0], ndim=3)),
        ("expsq_3d_general", K.ExpSquaredKernel([[1.0, 0.1, 0.2], [0.1, 2.0, 0.3], [0.2, 0.3, 1.5]], ndim=3)),
        ("sum_expsq_expsine2", 1.0 * K.ExpSquaredKernel(1.0, ndim=3)
         + 0.5 * K.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0), ndim=3, axes=1)),
        ("ratquad", K.RationalQuadraticKernel(log_alpha=0.3, metric=1.2, ndim=3)),
        ("exp_axes", K.ExpKernel(1.2, ndim=3, axes=[0, 2])),
        ("cos_x_localgauss", K.CosineKernel(log_period=0.5, ndim=3, axes=0)
         * K.LocalGaussianKernel(location=0.1, log_width=0.2, ndim=3, axes=1)),
        ("poly_lin_dot", K.PolynomialKernel(log_sigma2=0.1, order=3, ndim=3)
         + K.LinearKernel(log_gamma2=0.2, order=2, ndim=3) + K.DotProductKernel(ndim=3)),
        ("expsq_block", K.ExpSquaredKernel(1.0, ndim=3, block=[(-0.5, 0.5)] * 3)),
        ("cfg5_1d", 1.0 * K.ExpSquaredKernel(1.0) + 0.5 * K.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0))),
        ("const_plus", 0.3 + K.Matern32Kernel(2.0, ndim=2)),
        ("empty", K.EmptyKernel(ndim=2) + K.ConstantKernel(log_constant=0.1, ndim=2)),
    ]
    return zoo


def reference_kernel_list():
    """The kernel instances the reference's own suite exercises (tests/test_kernels.py:19-64 `kernels_to_test` and the
    `test_stationary` variants :83-128), built from OUR classes."""
    from george_b200 import kernels
    out = [
        kernels.ConstantKernel(log_constant=0.1),
        kernels.ConstantKernel(log_constant=10.0, ndim=2),
        kernels.ConstantKernel(log_constant=5.0, ndim=5),
        kernels.DotProductKernel(),
        kernels.DotProductKernel(ndim=2),
        kernels.DotProductKernel(ndim=5, axes=0),
        kernels.CosineKernel(log_period=1.0),
        kernels.CosineKernel(log_period=0.5, ndim=2),
        kernels.CosineKernel(log_period=0.5, ndim=2, axes=1),
        kernels.CosineKernel(log_period=0.75, ndim=5, axes=[2, 3]),
        kernels.ExpSine2Kernel(gamma=0.4, log_period=1.0),
        kernels.ExpSine2Kernel(gamma=12., log_period=0.5, ndim=2),
        kernels.ExpSine2Kernel(gamma=17., log_period=0.5, ndim=2, axes=1),
        kernels.ExpSine2Kernel(gamma=13.7, log_period=-0.75, ndim=5, axes=[2, 3]),
        kernels.ExpSine2Kernel(gamma=-0.7, log_period=0.75, ndim=5, axes=[2, 3]),
        kernels.ExpSine2Kernel(gamma=-10, log_period=0.75),
        kernels.LocalGaussianKernel(log_width=0.5, location=1.0),
        kernels.LocalGaussianKernel(log_width=0.1, location=0.5, ndim=2),
        kernels.LocalGaussianKernel(log_width=1.5, location=-0.5, ndim=2, axes=1),
        kernels.LocalGaussianKernel(log_width=2.0, location=0.75, ndim=5, axes=[2, 3]),
        kernels.LinearKernel(order=0, log_gamma2=0.0),
        kernels.LinearKernel(order=2, log_gamma2=0.0),
        kernels.LinearKernel(order=5, log_gamma2=1.0, ndim=2),
        kernels.LinearKernel(order=3, log_gamma2=-1.0, ndim=5, axes=2),
        kernels.LinearKernel(order=0, log_gamma2=0.0) + kernels.LinearKernel(order=1, log_gamma2=-1.0)
        + kernels.LinearKernel(order=2, log_gamma2=-2.0),
        kernels.PolynomialKernel(order=0, log_sigma2=-10.0),
        kernels.PolynomialKernel(order=2, log_sigma2=-10.0),
        kernels.PolynomialKernel(order=2, log_sigma2=0.0),
        kernels.PolynomialKernel(order=5, log_sigma2=1.0, ndim=2),
        kernels.PolynomialKernel(order=3, log_sigma2=-1.0, ndim=5, axes=2),
        12. * kernels.ExpSine2Kernel(gamma=0.4, log_period=1.0, ndim=5),
        12. * kernels.ExpSquaredKernel(0.4, ndim=3) + 0.1,
    ]
    stationary = [
        (kernels.ExpKernel, {}), (kernels.ExpSquaredKernel, {}), (kernels.Matern32Kernel, {}),
        (kernels.Matern52Kernel, {}), (kernels.RationalQuadraticKernel, dict(log_alpha=np.log(1.0))),
        (kernels.RationalQuadraticKernel, dict(log_alpha=np.log(0.1))),
        (kernels.RationalQuadraticKernel, dict(log_alpha=np.log(10.0))),
    ]
    for cls, kw in stationary:
        out += [cls(metric=0.1, **kw), cls(metric=1.0, **kw), cls(metric=10.0, **kw),
                cls(metric=[1.0, 0.1, 10.0], ndim=3, **kw), cls(metric=1.0, ndim=3, **kw),
                cls(metric=1.0, ndim=3, axes=2, **kw), cls(metric=1.0, ndim=3, axes=2, block=(-0.1, 0.1), **kw)]
    return out

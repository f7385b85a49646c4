# -*- coding: utf-8 -*-
"""Host-side logic that needs no GPU: modeling protocol, kernel specs, flattening, GP bookkeeping
(modelled on the reference's tests/test_modeling.py, test_kernels.py construction cases, test_metrics.py)."""
import pickle

import numpy as np
import pytest


def test_model_protocol():
    from george_b200.modeling import Model, ModelSet, ConstantModel, CallableModel

    class Line(Model):
        parameter_names = ("m", "b")

        def get_value(self, x):
            return self.m * x + self.b

    ln = Line(m=2.0, b=1.0, bounds={"m": (0, 5)})
    assert ln.get_parameter_names() == ("m", "b") and len(ln) == 2
    assert np.allclose(ln.get_value(np.array([1.0, 2.0])), [3.0, 5.0])
    ln.freeze_parameter("b")
    assert ln.get_parameter_names() == ("m",) and ln.vector_size == 1
    ln.set_parameter_vector([3.0])
    assert ln.m == 3.0 and ln.dirty
    assert ln.log_prior() == 0.0
    assert not ln.check_parameter_vector([9.0])
    ln["m"] = 4.0
    assert ln[0] == 4.0
    g = ln.get_gradient(np.array([1.0, 2.0]))
    assert g.shape == (1, 2) and np.allclose(g[0], [1.0, 2.0], atol=1e-4)
    with pytest.raises(ValueError):
        Line(m=9.0, b=0.0, bounds={"m": (0, 5)})
    with pytest.raises(ValueError):
        Line(m=1.0)
    ms = ModelSet([("a", Line(1.0, 2.0)), ("c", ConstantModel(3.0))])
    assert ms.get_parameter_names() == ("a:m", "a:b", "c:value")
    ms.set_parameter("a:b", 7.0)
    assert ms.get_parameter("a:b") == 7.0
    ms.freeze_parameter("c:value")
    assert ms.vector_size == 2
    assert np.allclose(CallableModel(np.sin).get_value(np.array([0.0])), [0.0])


def test_kernel_construction_and_arithmetic():
    from george_b200 import kernels as K
    k = 3.0 * K.ExpSquaredKernel(2.0, ndim=2)
    assert k.k1.kernel_type == 8 and np.allclose(k.k1.log_constant, np.log(3.0 / 2))
    assert k.k2.kernel_type == 9 and k.k2.metric.metric_type == 0
    assert np.allclose(k.k2.metric.get_parameter_vector(), [np.log(2.0)])
    k2 = K.Matern32Kernel([1.0, 4.0], ndim=3, axes=[0, 2])
    assert k2.metric.metric_type == 1 and list(k2.axes) == [0, 2] and not k2.blocked
    k3 = K.ExpSquaredKernel([[2.0, 0.5], [0.5, 1.0]], ndim=2)
    assert k3.metric.metric_type == 2 and len(k3.metric.get_parameter_vector()) == 3
    assert np.allclose(k3.metric.to_matrix(), [[2.0, 0.5], [0.5, 1.0]])
    kb = K.ExpSquaredKernel(1.0, ndim=2, block=[(0, 1), (-1, 1)])
    assert kb.blocked and kb.block == [(0.0, 1.0), (-1.0, 1.0)]
    with pytest.raises(ValueError):
        K.ExpSquaredKernel()
    with pytest.raises(ValueError):
        K.LinearKernel(log_gamma2=0.0)
    with pytest.raises(ValueError):
        K.ExpSquaredKernel(1.0) + K.ExpSquaredKernel(1.0, ndim=2)
    s = K.ExpSine2Kernel(gamma=1.0, log_period=0.5) + 2.0
    assert s.k1.kernel_type == 8 and s.k2.kernel_type == 7
    assert (np.float64(2.0) * K.ExpKernel(1.0)).k1.kernel_type == 8
    k.set_parameter_vector([0.1, 0.2])
    assert k.dirty and np.allclose(k.get_parameter_vector(), [0.1, 0.2])
    kk = pickle.loads(pickle.dumps(k))
    assert np.allclose(kk.get_parameter_vector(), [0.1, 0.2])
    assert "ExpSquaredKernel" in repr(k)


def test_flatten_program():
    from george_b200 import kernels as K
    from george_b200._spec import flatten, num_params
    k = 1.0 * K.ExpSquaredKernel(1.0, ndim=3) + 0.5 * K.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0), ndim=3, axes=1)
    spec = flatten(k)
    ops = [spec.nodes[i].op for i in range(spec.n_nodes)]
    assert ops == [0, 0, 2, 0, 0, 2, 1]  # const expsq * const expsine2 * +
    assert spec.ndim == 3 and num_params(spec) == len(k) == 5
    assert spec.nodes[4].kernel_type == 7 and spec.nodes[4].naxes == 1 and spec.nodes[4].axes[0] == 1
    assert np.allclose(list(spec.nodes[4].params)[:2], [1.0, np.log(3.0)])

    class Bogus(object):
        pass
    with pytest.raises(ValueError):
        flatten(Bogus())


def test_gp_bookkeeping_without_device():
    import george_b200 as george
    from george_b200 import kernels
    gp = george.GP(2.0 * kernels.Matern32Kernel(1.0), mean=1.5, white_noise=-3.0, fit_white_noise=True,
                   solver=george.HODLRSolver, tol=1e-8, min_size=64)
    assert gp.solver_type is george.HODLRSolver and gp.solver_kwargs == {"tol": 1e-8, "min_size": 64}
    assert gp.get_parameter_names() == ("white_noise:value", "kernel:k1:log_constant", "kernel:k2:metric:log_M_0_0")
    assert george.GP().solver_type is george.TrivialSolver
    with pytest.raises(RuntimeError):
        gp.recompute()
    with pytest.raises(ValueError):
        gp.parse_samples(np.zeros((3, 2)))
    t = george.TrivialSolver()
    t.compute(np.zeros((3, 1)), np.array([1.0, 2.0, 4.0]))
    assert np.allclose(t.log_determinant, 2 * np.log(8.0))
    assert np.allclose(t.dot_solve(np.ones(3)), 1 + 0.25 + 1 / 16.0)
    gpt = george.GP()
    gpt.compute(np.arange(3.0), np.array([1.0, 2.0, 4.0]))
    assert np.isfinite(gpt.log_likelihood(np.ones(3)))
    h = george.HODLRSolver(kernels.ExpSquaredKernel(1.0))
    assert (h.min_size, h.tol, h.seed) == (100, 0.1, 42)  # reference defaults (solvers/hodlr.py:43)
    with pytest.raises(NotImplementedError):
        h.apply_sqrt(np.zeros(3))
    st = pickle.loads(pickle.dumps(h))
    assert not st.computed


def test_matrix_free_entry_points_validate_shapes_before_touching_the_device():
    """KernelInterface.matvec / gradient_contract: argument checks are host-side (same exception types as the rest of
    the interface, kernel_interface.cpp:51); with valid arguments and no GPU the call must raise, never compute."""
    from george_b200 import kernels as K, _lib
    from george_b200._spec import DimensionMismatch
    k = 1.0 * K.Matern32Kernel(1.0, ndim=2)
    ki = k.kernel
    x1, x2 = np.zeros((5, 2)), np.zeros((7, 2))
    with pytest.raises(DimensionMismatch):
        ki.matvec(x1, x2, np.zeros(6))                      # v does not match x2
    with pytest.raises(DimensionMismatch):
        ki.matvec(x1, np.zeros((7, 3)), np.zeros(7))        # wrong input dimension
    with pytest.raises(DimensionMismatch):
        ki.matvec(x1, x2, np.zeros(7), diag=np.zeros(5))    # a diagonal term needs a square operator
    with pytest.raises(ValueError):
        ki.matvec(np.zeros(5), x2, np.zeros(7))             # 1-D coordinates
    with pytest.raises(DimensionMismatch):
        ki.gradient_contract(np.ones(ki.size, dtype=np.uint32), x1, np.zeros((5, 4)))
    with pytest.raises(DimensionMismatch):
        ki.gradient_contract(np.ones(ki.size + 1, dtype=np.uint32), x1, np.zeros((5, 5)))
    if _lib.load().bgp_device_count() == 0:
        with pytest.raises(_lib.BGPError):
            ki.matvec(x1, x2, np.zeros(7))
        with pytest.raises(_lib.BGPError):
            ki.gradient_contract(np.ones(ki.size, dtype=np.uint32), x1, np.zeros((5, 5)))


def test_grad_log_likelihood_uses_solver_grad_terms_and_falls_back():
    """GP.grad_log_likelihood composes the gradient from `solver.grad_terms` when the plug-in offers it, and from
    `get_inverse` + `KernelInterface.gradient_contract` otherwise (reference gp.py:406-468 semantics either way)."""
    import george_b200 as george
    from george_b200 import kernels as K

    calls = []

    class FakeSolver(object):
        def __init__(self, kernel, **kw):
            self.kernel = kernel
            self.computed = False
            self.log_determinant = 0.0

        def compute(self, x, yerr):
            self.n = len(x)
            self.computed = True

        def apply_inverse(self, y, in_place=False):
            return np.array(y, dtype=float)

        def grad_terms(self, r, which):
            calls.append(("grad_terms", which.copy()))
            return np.asarray(r, dtype=float), np.arange(1.0, which.size + 1.0), np.full(self.n, 2.0)

    kernel = 2.0 * K.ExpSquaredKernel(1.0)
    kernel.freeze_parameter("k1:log_constant")
    gp = george.GP(kernel, solver=FakeSolver, white_noise=np.log(0.1), fit_white_noise=True, mean=0.5, fit_mean=True)
    x = np.linspace(0, 1, 6)
    y = np.sin(x)
    gp.compute(x, 0.1)
    g = gp.grad_log_likelihood(y)
    assert calls and list(calls[0][1]) == [0, 1]            # the frozen constant is masked out
    names = gp.get_parameter_names()
    assert len(g) == len(names) == 3
    assert np.isclose(g[0], np.sum(y - 0.5))                # mean: dmu . alpha with alpha = r
    assert np.isclose(g[1], 0.5 * 0.1 * 2.0 * 6)            # white noise: 0.5 sum(exp(wn) diagA)
    assert np.isclose(g[2], 0.5 * 2.0)                      # kernel: 0.5 * g[mask] -> entry 1 of (1, 2)


def test_user_kernel_codegen_is_up_to_date_and_wired():
    """kernels/*.yml -> csrc/user_kernels.cuh + user_kernels.py (tools/generate_kernels.py, the counterpart of the
    reference's generate_kernels.py:10-42): the checked-in files are what the YAML produces, the classes exist with the
    reference's constructor conventions, and the flattened program is accepted by the library's host-side validation."""
    import os
    import subprocess
    import sys
    import ctypes as C
    from george_b200 import kernels as K, _lib
    from george_b200._spec import flatten
    from george_b200.user_kernels import USER_KERNEL_TABLE, BGP_K_USER0
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "generate_kernels.py"), "--check"],
                          stdout=subprocess.DEVNULL)
    names = [r[0] for r in USER_KERNEL_TABLE]
    assert names == ["CauchyKernel", "DampedCosineKernel"]
    k1 = K.CauchyKernel(metric=2.0)                       # stationary: metric / ndim / axes / block like every built-in
    assert k1.stationary and k1.kernel_type == BGP_K_USER0 and k1.get_parameter_names() == ("metric:log_M_0_0",)
    k2 = K.DampedCosineKernel(log_period=0.3, log_decay=1.0, ndim=2, axes=1)
    assert not k2.stationary and k2.kernel_type == BGP_K_USER0 + 1
    assert k2.get_parameter_names() == ("log_period", "log_decay")
    lib = _lib.load()
    for kern, npar in ((k1, 1), (k2, 2), (1.5 * k1 + K.Matern32Kernel(1.0) * K.CauchyKernel(0.5), 4)):
        spec = flatten(kern)
        assert lib.bgp_spec_validate(C.byref(spec)) == 0
        n = C.c_int()
        assert lib.bgp_spec_num_params(C.byref(spec), C.byref(n)) == 0 and n.value == npar
    bad = flatten(k1)
    bad.nodes[0].kernel_type = BGP_K_USER0 + len(names)   # one past the last compiled-in kernel
    assert lib.bgp_spec_validate(C.byref(bad)) != 0

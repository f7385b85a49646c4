# -*- coding: utf-8 -*-
"""
Kernel *specification* objects — the ``george.kernels`` surface.

A kernel here is only a tree of attribute-bearing objects (the reference's are generated from
``kernels/*.yml`` into ``src/george/kernels.py``); all numeric work is done on the B200 by the fused
kernel-matrix build in ``csrc/`` after ``_spec.flatten`` has turned the tree into the POD program of
``include/bgp.h``.  The attribute protocol is exactly the one ``src/george/include/george/parser.h:14-509``
reads (``is_kernel``, ``kernel_type``, ``operator_type``, ``k1``/``k2``, ``metric``, ``blocked``,
``min_block``/``max_block``, ``ndim``, ``axes`` and the per-kernel parameters), so these objects can be handed
unchanged to the reference's own compiled ``kernel_interface`` — which is how the parity tests cross-check them.

Kernel ids, parameter names and constructor signatures follow the reference (``kernels.py:273-966``);
scalar arithmetic follows ``kernels.py:83-100``: ``c * k`` becomes ``ConstantKernel(log(c / ndim)) * k``
because the constant kernel is summed over its axes (``kernels.h:1720-1732``).
"""

import sys

import numpy as np

from .kernel_interface import KernelInterface
from .metrics import Metric, Subspace
from .modeling import Model, ModelSet

__all__ = [
    "Kernel", "Sum", "Product",
    "LinearKernel", "RationalQuadraticKernel", "ExpKernel", "LocalGaussianKernel", "EmptyKernel",
    "CosineKernel", "Matern52Kernel", "ExpSine2Kernel", "ConstantKernel", "ExpSquaredKernel",
    "Matern32Kernel", "PolynomialKernel", "DotProductKernel",
]


class Kernel(ModelSet):
    """Abstract kernel: every concrete kernel and both operators derive from this."""

    is_kernel = True
    kernel_type = -1

    # numpy scalars on the left of ``*`` / ``+`` would otherwise broadcast over the kernel object
    __array_priority__ = np.inf

    def __array_wrap__(self, array, context=None):
        if context is None:
            raise TypeError("Invalid operation")
        ufunc, args, _ = context
        if ufunc.__name__ == "multiply":
            return float(args[0]) * args[1]
        if ufunc.__name__ == "add":
            return float(args[0]) + args[1]
        raise TypeError("Invalid operation")

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_kernel"] = None  # never pickle a native handle (reference kernels.py:52-55)
        return state

    def __getattr__(self, name):
        # parameters of the un-named base model are visible as attributes of the kernel itself
        if "models" in self.__dict__:
            if name in self.models:
                return self.models[name]
            if None in self.models:
                return getattr(self.models[None], name)
        raise AttributeError(name)

    @property
    def kernel(self):
        """A fresh evaluation interface holding a snapshot of the current parameters (kernels.py:67-69)."""
        return KernelInterface(self)

    def __repr__(self):
        base = self.models[None]
        parts = ["{0}={1}".format(k, getattr(base, k)) for k in base.parameter_names]
        if self.stationary:
            parts += ["metric={0}".format(repr(self.metric)), "block={0}".format(repr(self.block))]
        else:
            parts += ["ndim={0}".format(self.ndim), "axes={0}".format(repr(self.axes))]
        return "{0}({1})".format(self.__class__.__name__, ", ".join(parts))

    # -- arithmetic -------------------------------------------------------------------------------------------
    def _as_constant(self, b):
        return ConstantKernel(log_constant=np.log(float(b) / self.ndim), ndim=self.ndim)

    def __add__(self, b):
        if not hasattr(b, "is_kernel"):
            return Sum(self._as_constant(b), self)
        return Sum(self, b)

    def __radd__(self, b):
        return self.__add__(b)

    def __mul__(self, b):
        if not hasattr(b, "is_kernel"):
            return Product(self._as_constant(b), self)
        return Product(self, b)

    def __rmul__(self, b):
        return self.__mul__(b)

    # -- evaluation (delegates to the device through KernelInterface) -----------------------------------------
    def get_value(self, x1, x2=None, diag=False):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        if x2 is None:
            if diag:
                return self.kernel.value_diagonal(x1, x1)
            return self.kernel.value_symmetric(x1)
        x2 = np.ascontiguousarray(x2, dtype=np.float64)
        if diag:
            return self.kernel.value_diagonal(x1, x2)
        return self.kernel.value_general(x1, x2)

    def matvec(self, x1, x2, v, diag=None):
        """``get_value(x1, x2) @ v`` evaluated matrix-free on the device (``KernelInterface.matvec``)."""
        return self.kernel.matvec(np.ascontiguousarray(x1, dtype=np.float64),
                                  np.ascontiguousarray(x2, dtype=np.float64), v, diag=diag)

    def get_gradient(self, x1, x2=None, include_frozen=False):
        mask = np.ones(self.full_size, dtype=bool) if include_frozen else self.unfrozen_mask
        which = mask.astype(np.uint32)
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        if x2 is None:
            g = self.kernel.gradient_symmetric(which, x1)
        else:
            x2 = np.ascontiguousarray(x2, dtype=np.float64)
            g = self.kernel.gradient_general(which, x1, x2)
        return g[:, :, mask]

    def get_x1_gradient(self, x1, x2=None):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        x2 = x1 if x2 is None else np.ascontiguousarray(x2, dtype=np.float64)
        return self.kernel.x1_gradient_general(x1, x2)

    def get_x2_gradient(self, x1, x2=None):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        x2 = x1 if x2 is None else np.ascontiguousarray(x2, dtype=np.float64)
        return self.kernel.x2_gradient_general(x1, x2)

    def test_gradient(self, x1, x2=None, eps=1.32e-6, **kwargs):
        """Centred finite-difference check of the hyper-parameter gradient (reference kernels.py:145-165)."""
        p = self.get_parameter_vector()
        analytic = self.get_gradient(x1, x2=x2)
        for i, pi in enumerate(p):
            p[i] = pi + eps
            self.set_parameter_vector(p)
            plus = self.get_value(x1, x2=x2)
            p[i] = pi - eps
            self.set_parameter_vector(p)
            minus = self.get_value(x1, x2=x2)
            p[i] = pi
            self.set_parameter_vector(p)
            fd = 0.5 * (plus - minus) / eps
            assert np.allclose(analytic[:, :, i], fd, **kwargs), \
                "incorrect gradient for parameter '{0}' ({1})".format(self.get_parameter_names()[i], i)


    def _test_x_gradient(self, side, x1, x2, eps, kwargs):
        kwargs["atol"] = kwargs.get("atol", 0.5 * eps)
        analytic = (self.get_x1_gradient if side == 1 else self.get_x2_gradient)(x1, x2=x2)
        pts = [np.array(x1, dtype=np.float64), np.array(x1 if x2 is None else x2, dtype=np.float64)]
        moved = pts[side - 1]
        for i in range(len(moved)):
            for k in range(self.ndim):
                x0 = moved[i, k]
                moved[i, k] = x0 + eps
                plus = self.get_value(pts[0], x2=pts[1])
                moved[i, k] = x0 - eps
                minus = self.get_value(pts[0], x2=pts[1])
                moved[i, k] = x0
                fd = 0.5 * (plus - minus) / eps
                if side == 1:
                    assert np.allclose(analytic[i, :, k], fd[i], **kwargs)
                else:
                    assert np.allclose(analytic[:, i, k], fd[:, i], **kwargs)

    def test_x1_gradient(self, x1, x2=None, eps=1.32e-6, **kwargs):
        """Centred finite-difference check of d k / d x1 (reference kernels.py:166-182)."""
        self._test_x_gradient(1, x1, x2, eps, kwargs)

    def test_x2_gradient(self, x1, x2=None, eps=1.32e-6, **kwargs):
        """Centred finite-difference check of d k / d x2 (reference kernels.py:184-200)."""
        self._test_x_gradient(2, x1, x2, eps, kwargs)


class _operator(Kernel):
    is_kernel = False
    kernel_type = -1
    operator_type = -1

    def __init__(self, k1, k2):
        if k1.ndim != k2.ndim:
            raise ValueError("Dimension mismatch")
        self.ndim = k1.ndim
        self._dirty = True
        super(_operator, self).__init__([("k1", k1), ("k2", k2)])

    @property
    def k1(self):
        return self.models["k1"]

    @property
    def k2(self):
        return self.models["k2"]

    @property
    def dirty(self):
        return self._dirty or self.k1.dirty or self.k2.dirty

    @dirty.setter
    def dirty(self, v):
        self._dirty = v
        self.k1.dirty = False
        self.k2.dirty = False


class Sum(_operator):
    is_kernel = False
    operator_type = 0

    def __repr__(self):
        return "{0} + {1}".format(self.k1, self.k2)


class Product(_operator):
    is_kernel = False
    operator_type = 1

    def __repr__(self):
        return "{0} * {1}".format(self.k1, self.k2)


# ---------------------------------------------------------------------------------------------------------------
# Concrete kernels.  One table row per kernels/*.yml of the reference:
#   (class name, kernel_type id, stationary, hyper-parameters, constants, one-line formula)
# kernel_type is the reference's enumeration (kernels.py:273,322,391,464,500,546,588,666,712,753,822,897,944).
# ---------------------------------------------------------------------------------------------------------------
_KERNEL_TABLE = [
    ("LinearKernel", 0, False, ("log_gamma2",), ("order",), "k = (x_i . x_j)^P / gamma^2"),
    ("RationalQuadraticKernel", 1, True, ("log_alpha",), (), "k(r2) = [1 + r2 / (2 alpha)]^(-alpha)"),
    ("ExpKernel", 2, True, (), (), "k(r2) = exp(-sqrt(r2))"),
    ("LocalGaussianKernel", 3, False, ("location", "log_width"), (),
     "k = exp(-[(x_i - x0)^2 + (x_j - x0)^2] / (2 w))"),
    ("EmptyKernel", 4, False, (), (), "k = 0"),
    ("CosineKernel", 5, False, ("log_period",), (), "k = cos(2 pi |x_i - x_j| / P)"),
    ("Matern52Kernel", 6, True, (), (), "k(r2) = (1 + sqrt(5 r2) + 5 r2 / 3) exp(-sqrt(5 r2))"),
    ("ExpSine2Kernel", 7, False, ("gamma", "log_period"), (), "k = exp(-Gamma sin^2(pi |x_i - x_j| / P))"),
    ("ConstantKernel", 8, False, ("log_constant",), (), "k = c (per axis)"),
    ("ExpSquaredKernel", 9, True, (), (), "k(r2) = exp(-r2 / 2)"),
    ("Matern32Kernel", 10, True, (), (), "k(r2) = (1 + sqrt(3 r2)) exp(-sqrt(3 r2))"),
    ("PolynomialKernel", 11, False, ("log_sigma2",), ("order",), "k = (x_i . x_j + sigma^2)^P"),
    ("DotProductKernel", 12, False, (), (), "k = x_i . x_j"),
]


def _block_get(self):
    if not self.blocked:
        return None
    return list(zip(self.min_block, self.max_block))


def _block_set(self, block):
    naxes = len(self.axes)
    if block is None:
        self.blocked = False
        self.min_block = np.full(naxes, -np.inf)
        self.max_block = np.full(naxes, np.inf)
        return
    block = np.atleast_2d(block)
    if block.shape != (naxes, 2):
        raise ValueError("dimension mismatch in block specification")
    self.blocked = True
    self.min_block, self.max_block = np.array(block[:, 0]), np.array(block[:, 1])


def _make_kernel_class(name, kernel_type, stationary, params, constants, formula):
    base_cls = type("Base" + name, (Model,), {"parameter_names": tuple(params), "__module__": __name__})

    if stationary:
        def __init__(self, *args, **kwargs):
            order = list(params) + ["metric", "metric_bounds", "lower", "block", "bounds", "ndim", "axes"]
            opts = dict(metric=None, metric_bounds=None, lower=True, block=None, bounds=None, ndim=1, axes=None)
            opts.update(_bind(name, order, args, kwargs))
            if opts["metric"] is None:
                raise ValueError("missing required parameter 'metric'")
            metric = Metric(opts["metric"], bounds=opts["metric_bounds"], ndim=opts["ndim"], axes=opts["axes"],
                            lower=opts["lower"])
            self.ndim, self.axes = metric.ndim, metric.axes
            self.block = opts["block"]
            base_kwargs = {k: opts.get(k) for k in params}
            if opts["bounds"] is not None:
                base_kwargs["bounds"] = opts["bounds"]
            Kernel.__init__(self, [(None, base_cls(**base_kwargs)), ("metric", metric)])
            self.dirty = True
        ns = {"block": property(_block_get, _block_set)}
    else:
        def __init__(self, *args, **kwargs):
            order = list(params) + list(constants) + ["bounds", "ndim", "axes"]
            opts = dict(bounds=None, ndim=1, axes=None)
            opts.update(_bind(name, order, args, kwargs))
            for c in constants:
                if opts.get(c) is None:
                    raise ValueError("missing required parameter '{0}'".format(c))
                setattr(self, c, opts[c])
            self.subspace = Subspace(opts["ndim"], axes=opts["axes"])
            self.ndim, self.axes = self.subspace.ndim, self.subspace.axes
            base_kwargs = {k: opts.get(k) for k in params}
            if opts["bounds"] is not None:
                base_kwargs["bounds"] = opts["bounds"]
            Kernel.__init__(self, [(None, base_cls(**base_kwargs))])
            self.dirty = True
        ns = {}

    ns.update({
        "__init__": __init__, "__module__": __name__, "__doc__": formula,
        "kernel_type": kernel_type, "stationary": stationary,
    })
    return base_cls, type(name, (Kernel,), ns)


def _bind(cls_name, order, args, kwargs):
    """Positional/keyword binding in the reference's argument order."""
    if len(args) > len(order):
        raise TypeError("{0}() takes at most {1} arguments".format(cls_name, len(order)))
    bound = dict(zip(order, args))
    for k, v in kwargs.items():
        if k not in order:
            raise TypeError("{0}() got an unexpected keyword argument '{1}'".format(cls_name, k))
        if k in bound:
            raise TypeError("{0}() got multiple values for argument '{1}'".format(cls_name, k))
        bound[k] = v
    return bound


# user kernels: rows generated from kernels/*.yml by tools/generate_kernels.py (the reference generates its whole
# kernels.py this way, generate_kernels.py:10-42); the device functors are in csrc/user_kernels.cuh
try:
    from .user_kernels import USER_KERNEL_TABLE
except ImportError:  # pragma: no cover
    USER_KERNEL_TABLE = []
_KERNEL_TABLE = _KERNEL_TABLE + [tuple(r) for r in USER_KERNEL_TABLE]
__all__ += [r[0] for r in USER_KERNEL_TABLE]

_module = sys.modules[__name__]
for _row in _KERNEL_TABLE:
    _base, _cls = _make_kernel_class(*_row)
    setattr(_module, _base.__name__, _base)
    setattr(_module, _cls.__name__, _cls)
del _row, _base, _cls

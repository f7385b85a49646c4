# -*- coding: utf-8 -*-
"""
``BasicSolver`` — dense Cholesky on the B200.

Drop-in for the reference's ``src/george/solvers/basic.py:11-121`` (``kernel.get_value`` + ``scipy.linalg.cholesky``
/ ``cho_solve``): same constructor, methods, properties and return shapes, but the kernel matrix is generated,
factorised and solved against entirely on the device (``csrc/kmat.cu`` + ``csrc/dense.cu``) and never visits the host.
A non positive-definite matrix raises ``numpy.linalg.LinAlgError`` exactly like scipy, which ``GP.recompute`` relies on.
"""

import ctypes as C

import numpy as np

from .. import _lib
from .._spec import flatten

__all__ = ["BasicSolver"]


class _DenseHandle(object):
    """Owns a ``bgp_dense_t*``; never pickled."""

    def __init__(self):
        self.lib = _lib.load()
        self.ptr = C.c_void_p()
        _lib.check(self.lib.bgp_dense_create(C.byref(self.ptr)))

    def __del__(self):
        if getattr(self, "ptr", None) is not None and self.ptr:
            self.lib.bgp_dense_destroy(self.ptr)
            self.ptr = None


class BasicSolver(object):

    def __init__(self, kernel):
        self.kernel = kernel
        self._computed = False
        self._log_det = None
        self._handle = None
        self._n = 0

    @property
    def computed(self):
        """Has the covariance matrix been built and factorised (by :func:`compute`)?"""
        return self._computed

    @computed.setter
    def computed(self, v):
        self._computed = v

    @property
    def log_determinant(self):
        """log|K|; ``None`` before :func:`compute`."""
        return self._log_det

    @log_determinant.setter
    def log_determinant(self, v):
        self._log_det = v

    def compute(self, x, yerr):
        """Build K(x, x) + diag(yerr^2) on the device and Cholesky-factorise it (basic.py:51-70)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        n, ndim = x.shape
        yerr = np.ascontiguousarray(np.broadcast_to(np.asarray(yerr, dtype=np.float64), (n,)))
        spec = flatten(self.kernel)
        if self._handle is None:
            self._handle = _DenseHandle()
        lib = self._handle.lib
        self._computed = False
        _lib.check(lib.bgp_dense_compute(self._handle.ptr, C.byref(spec), _lib.ptr(x), n, ndim, _lib.ptr(yerr)))
        ld = C.c_double()
        _lib.check(lib.bgp_dense_log_determinant(self._handle.ptr, C.byref(ld)))
        self._n = n
        self._has_inputs = True
        self.log_determinant = ld.value
        self.computed = True

    def _require(self):
        if self._handle is None or not self._computed:
            raise RuntimeError("you must call 'compute' first")

    def apply_inverse(self, y, in_place=False):
        r"""Solve :math:`K\,x = y` for ``y`` of shape ``(n,)`` or ``(n, nrhs)`` (basic.py:72-87)."""
        self._require()
        y = np.asarray(y)
        if in_place and y.dtype == np.float64 and y.flags.f_contiguous and y.flags.writeable:
            b = y
        else:
            b = np.array(y, dtype=np.float64, order="F")
        if b.shape[0] != self._n:
            raise ValueError("dimension mismatch")
        nrhs = 1 if b.ndim == 1 else int(np.prod(b.shape[1:]))
        _lib.check(self._handle.lib.bgp_dense_apply_inverse(self._handle.ptr, _lib.ptr(b), nrhs, self._n))
        if in_place and b is not y:
            y[...] = b
            return y
        return b

    def dot_solve(self, y):
        r"""``y^T K^{-1} y`` (basic.py:89-102)."""
        self._require()
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.shape != (self._n,):
            raise ValueError("dimension mismatch")
        out = C.c_double()
        _lib.check(self._handle.lib.bgp_dense_dot_solve(self._handle.ptr, _lib.ptr(y), C.byref(out)))
        return out.value

    def apply_sqrt(self, r):
        """``r @ U`` with ``U`` the upper Cholesky factor (basic.py:104-114)."""
        self._require()
        r = np.ascontiguousarray(r, dtype=np.float64)
        r2 = r.reshape(-1, self._n)
        out = np.empty_like(r2)
        _lib.check(self._handle.lib.bgp_dense_apply_sqrt(self._handle.ptr, _lib.ptr(r2), r2.shape[0], _lib.ptr(out)))
        return out.reshape(r.shape)

    def get_inverse(self):
        """Dense ``K^{-1}`` (used by the gradient; basic.py:116-121)."""
        self._require()
        out = np.empty((self._n, self._n), dtype=np.float64)
        _lib.check(self._handle.lib.bgp_dense_get_inverse(self._handle.ptr, _lib.ptr(out)))
        return out

    def grad_terms(self, r, which):
        """``(alpha, g, diagA)`` for ``GP.grad_log_likelihood`` (gp.py:406-468), all computed on the device from the
        stored factor: ``alpha = K^-1 r``, ``g[p] = sum_ij (alpha alpha^T - K^-1)_ij dK_ij/dtheta_p`` over ALL kernel
        parameters (zeros where ``which`` is 0) and ``diagA = diag(alpha alpha^T - K^-1)``.  Replaces
        ``get_inverse()`` + ``kernel.get_gradient`` + ``einsum`` on the host.  Returns ``None`` for a solver restored from
        a pickle (its device handle holds the factor but neither kernel nor coordinates): the caller then composes the
        same quantities from ``get_inverse`` and ``KernelInterface.gradient_contract``."""
        self._require()
        if not getattr(self, "_has_inputs", True):
            return None
        r = np.ascontiguousarray(r, dtype=np.float64)
        if r.shape != (self._n,):
            raise ValueError("dimension mismatch")
        which = np.ascontiguousarray(which, dtype=np.uint32)
        alpha = np.empty(self._n, dtype=np.float64)
        g = np.zeros(max(which.size, 1), dtype=np.float64)
        diag = np.empty(self._n, dtype=np.float64)
        _lib.check(self._grad_terms_call(_lib.ptr(which), _lib.ptr(r), _lib.ptr(alpha), _lib.ptr(g), _lib.ptr(diag)))
        return alpha, g[:which.size], diag

    def _grad_terms_call(self, which, r, alpha, g, diag):
        return self._handle.lib.bgp_dense_grad_terms(self._handle.ptr, which, r, alpha, g, diag)

    # Device handles cannot be pickled.  Like the reference (which pickles its numpy factor, tests/test_pickle.py:21-36:
    # "Unpickled GP shouldn't need to be computed") the Cholesky factor travels with the pickle and is re-uploaded.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = None
        if self._handle is not None and self._computed:
            factor = np.empty((self._n, self._n), dtype=np.float64)
            _lib.check(self._handle.lib.bgp_dense_export_factor(self._handle.ptr, _lib.ptr(factor)))
            state["_pickled_factor"] = factor
        else:
            state["_computed"] = False
        return state

    def __setstate__(self, state):
        factor = state.pop("_pickled_factor", None)
        self.__dict__.update(state)
        self._handle = None
        if factor is not None:
            self._has_inputs = False
            try:
                self._handle = _DenseHandle()
                _lib.check(self._handle.lib.bgp_dense_import_factor(self._handle.ptr, _lib.ptr(factor), self._n,
                                                                   float(self._log_det)))
            except Exception:  # no device where it was unpickled: refactorise lazily on first use
                self._handle = None
                self._computed = False

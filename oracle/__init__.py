# -*- coding: utf-8 -*-
"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes binding of ``oracle/liboracle.so`` (the CPU restatement of the reference algorithm, see
``hodlr_oracle.cpp`` / ``kernels_oracle.h``) plus a loader for ``oracle/_ref/kernel_interface*.so`` — the
reference's own ``kernel_interface.cpp`` compiled from the sources where they lie (``oracle/Makefile``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  Nothing under ``george_b200/`` does.
"""

import ctypes as C
import glob
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(so)
            for f in ("hodlr_oracle.cpp", "kernels_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.oracle_num_params.restype = C.c_int
        L.oracle_num_params.argtypes = [vp]
        for name, args in {
            "oracle_value_general": [vp, vp, i64, vp, i64, vp],
            "oracle_value_symmetric": [vp, vp, i64, vp],
            "oracle_value_diagonal": [vp, vp, vp, i64, vp],
            "oracle_gradient_general": [vp, vp, vp, i64, vp, i64, vp],
            "oracle_x_gradient_general": [vp, C.c_int, vp, i64, vp, i64, vp],
        }.items():
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = args
        L.oracle_hodlr_compute.restype = vp
        L.oracle_hodlr_compute.argtypes = [vp, vp, i64, i32, vp, i32, dbl, i32, i32]
        L.oracle_hodlr_compute2.restype = vp
        L.oracle_hodlr_compute2.argtypes = [vp, vp, i64, i32, vp, i32, dbl, i32, i32, i32]
        L.oracle_hodlr_free.restype = None
        L.oracle_hodlr_free.argtypes = [vp]
        L.oracle_hodlr_log_determinant.restype = dbl
        L.oracle_hodlr_log_determinant.argtypes = [vp]
        L.oracle_hodlr_num_evals.restype = C.c_uint64
        L.oracle_hodlr_num_evals.argtypes = [vp]
        L.oracle_hodlr_apply_inverse.restype = None
        L.oracle_hodlr_apply_inverse.argtypes = [vp, vp, i64, i64]
        L.oracle_hodlr_dot_solve.restype = dbl
        L.oracle_hodlr_dot_solve.argtypes = [vp, vp]
        L.oracle_hodlr_num_nodes.restype = i64
        L.oracle_hodlr_num_nodes.argtypes = [vp]
        L.oracle_hodlr_node_info.restype = None
        L.oracle_hodlr_node_info.argtypes = [vp, vp]
        L.oracle_hodlr_node_pivots.restype = C.c_int
        L.oracle_hodlr_node_pivots.argtypes = [vp, i64, vp, vp]
        L.oracle_mt19937_words.restype = None
        L.oracle_mt19937_words.argtypes = [C.c_uint32, C.c_int, vp]
        L.oracle_uniform_ints.restype = None
        L.oracle_uniform_ints.argtypes = [C.c_uint32, C.c_int, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _c(x):
    return np.ascontiguousarray(x, dtype=np.float64)


# ---- kernel values -------------------------------------------------------------------------------------------------
def value_general(spec, x1, x2):
    x1, x2 = _c(x1), _c(x2)
    out = np.empty((len(x1), len(x2)))
    assert lib().oracle_value_general(C.byref(spec), _p(x1), len(x1), _p(x2), len(x2), _p(out)) == 0
    return out


def value_symmetric(spec, x):
    x = _c(x)
    out = np.empty((len(x), len(x)))
    assert lib().oracle_value_symmetric(C.byref(spec), _p(x), len(x), _p(out)) == 0
    return out


def value_diagonal(spec, x1, x2):
    x1, x2 = _c(x1), _c(x2)
    out = np.empty(len(x1))
    assert lib().oracle_value_diagonal(C.byref(spec), _p(x1), _p(x2), len(x1), _p(out)) == 0
    return out


def gradient_general(spec, which, x1, x2):
    x1, x2 = _c(x1), _c(x2)
    which = np.ascontiguousarray(which, dtype=np.uint32)
    npar = lib().oracle_num_params(C.byref(spec))
    out = np.zeros((len(x1), len(x2), npar))
    assert lib().oracle_gradient_general(C.byref(spec), _p(which), _p(x1), len(x1), _p(x2), len(x2), _p(out)) == 0
    return out


def x_gradient_general(spec, side, x1, x2):
    """side 1: d k / d x1, side 2: d k / d x2 -> (n1, n2, ndim)"""
    x1, x2 = _c(x1), _c(x2)
    out = np.zeros((len(x1), len(x2), x1.shape[1]))
    assert lib().oracle_x_gradient_general(C.byref(spec), int(side), _p(x1), len(x1), _p(x2), len(x2), _p(out)) == 0
    return out


# ---- HODLR ---------------------------------------------------------------------------------------------------------
NODE_FIELDS = ("start", "size", "half", "is_leaf", "parent", "direction", "depth", "rank", "rng_draws",
               "dense_fallback")


class HODLR(object):
    """The restated reference solver (``_hodlr.cpp:36-112``)."""

    def __init__(self, spec, x, yerr, min_size=100, tol=0.1, seed=42, rng_mode=1, exhaust=0):
        """rng_mode: 1 = the reference's single shared mt19937 (hodlr.h:35,58-61), 0 = one stream per node (the CUDA
        path's level-parallel mode).  exhaust: 0 = the reference's dense fallback (hodlr.h:161-176), 1 = keep the
        low-rank factors when every row has been rejected (the CUDA path's ``exhaust="lowrank"``)."""
        x = _c(x)
        if x.ndim == 1:
            x = x[:, None]
        yerr = _c(yerr)
        self.n = len(x)
        self._h = lib().oracle_hodlr_compute2(C.byref(spec), _p(x), len(x), x.shape[1], _p(yerr), int(min_size),
                                              float(tol), int(seed), int(rng_mode), int(exhaust))
        if not self._h:
            raise ValueError("oracle: invalid kernel program")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_hodlr_free(self._h)
            self._h = None

    @property
    def log_determinant(self):
        return lib().oracle_hodlr_log_determinant(self._h)

    @property
    def num_evals(self):
        return int(lib().oracle_hodlr_num_evals(self._h))

    def apply_inverse(self, y):
        y = np.array(y, dtype=np.float64)
        b = np.asfortranarray(y.reshape(self.n, -1))
        lib().oracle_hodlr_apply_inverse(self._h, _p(b), b.shape[1], self.n)
        return b.reshape(y.shape) if y.ndim > 1 else b[:, 0]

    def dot_solve(self, y):
        y = _c(y)
        return lib().oracle_hodlr_dot_solve(self._h, _p(y))

    def nodes(self):
        n = int(lib().oracle_hodlr_num_nodes(self._h))
        raw = np.zeros((n, len(NODE_FIELDS)), dtype=np.int32)
        lib().oracle_hodlr_node_info(self._h, _p(raw))
        return [dict(zip(NODE_FIELDS, (int(v) for v in row))) for row in raw]

    def pivots(self, node, rank):
        rows = np.zeros(max(rank, 1), dtype=np.int32)
        cols = np.zeros(max(rank, 1), dtype=np.int32)
        k = lib().oracle_hodlr_node_pivots(self._h, node, _p(rows), _p(cols))
        return rows[:k].copy(), cols[:k].copy()


def mt19937_words(seed, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().oracle_mt19937_words(seed, n, _p(out))
    return out


def uniform_ints(seed, sizes):
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    out = np.zeros(len(sizes), dtype=np.int32)
    lib().oracle_uniform_ints(seed, len(sizes), _p(sizes), _p(out))
    return out


# ---- the reference's own compiled kernel_interface (oracle/_ref) -------------------------------------------------
def reference_kernel_interface():
    """Import ``oracle/_ref/kernel_interface*.so`` (the reference's unmodified kernel_interface.cpp) or return None."""
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "kernel_interface*.so")))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("kernel_interface", hits[0])
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:  # ABI mismatch on a foreign box: treat as absent
        return None
    return mod

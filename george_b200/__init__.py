# -*- coding: utf-8 -*-
"""
george_b200 — a Blackwell-native (sm_100a) engine for the hot path of dfm/george:
``gp.compute()`` + ``gp.log_likelihood()`` (+ ``gp.predict()``), behind george's own ``GP`` / ``kernels`` /
solver-plugin surface.  Same names as ``george/__init__.py:3-18`` so ``import george_b200 as george`` is a drop-in.

The numeric work lives in ``george_b200/lib/libbgp_b200.so`` (hand-written CUDA, C ABI in ``include/bgp.h``); there is
no CPU fallback — importing works anywhere, computing requires a B200.
"""

__all__ = ["__version__", "kernels", "GP", "Metric", "TrivialSolver", "BasicSolver", "HODLRSolver"]

__version__ = "0.1.0"

from . import kernels
from .gp import GP
from .metrics import Metric
from .solvers import TrivialSolver, BasicSolver, HODLRSolver

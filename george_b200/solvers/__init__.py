# -*- coding: utf-8 -*-
"""Solver plugins (same names as the reference's ``src/george/solvers/__init__.py``)."""

__all__ = ["TrivialSolver", "BasicSolver", "HODLRSolver"]

from .trivial import TrivialSolver
from .basic import BasicSolver
from .hodlr import HODLRSolver

# -*- coding: utf-8 -*-
"""
``HODLRSolver`` — the O(N log^2 N) solver plugin, on the B200.

Drop-in for the reference's shim ``src/george/solvers/hodlr.py:12-76``: same constructor defaults
(``min_size=100, tol=0.1, seed=42``), ``compute`` builds a fresh native solver, ``apply_sqrt`` raises
``NotImplementedError``, pickling drops the native handle and clears ``computed`` so the GP refactorises lazily.
"""

from .basic import BasicSolver
from ._hodlr import HODLRSolver as HODLRSolverInterface

__all__ = ["HODLRSolver"]


class HODLRSolver(BasicSolver):

    def __init__(self, kernel, min_size=100, tol=0.1, seed=42, rng_mode=None, rank_capacity=0,
                 exhaust="dense"):
        # rng_mode=None: the reference's single shared mt19937 (bit-for-bit its pivot order) whenever the tolerance is
        # loose enough for the answer to DEPEND on the pivots (tol > 1e-6; the reference's default is 0.1), the
        # level-parallel per-node streams otherwise (the answer then agrees with the reference's far inside 1e-6)
        self.min_size = min_size
        self.tol = tol
        self.seed = seed
        self.rng_mode = rng_mode
        self.rank_capacity = rank_capacity
        self.exhaust = exhaust
        super(HODLRSolver, self).__init__(kernel)

    def compute(self, x, yerr):
        self.solver = HODLRSolverInterface()
        self.solver.compute(self.kernel, x, yerr, self.min_size, self.tol, self.seed, rng_mode=self.rng_mode,
                            rank_capacity=self.rank_capacity, exhaust=self.exhaust)
        self._log_det = self.solver.log_determinant
        self.computed = self.solver.computed

    def apply_inverse(self, y, in_place=False):
        return self.solver.apply_inverse(y, in_place=in_place)

    def dot_solve(self, y):
        return self.solver.dot_solve(y)

    def apply_sqrt(self, r):
        raise NotImplementedError("apply_sqrt is not implemented for the HODLRSolver")

    def get_inverse(self):
        return self.solver.get_inverse()

    def _require(self):
        if getattr(self, "solver", None) is None or not self._computed:
            raise RuntimeError("you must call 'compute' first")
        self._n = self.solver._n

    def _grad_terms_call(self, which, r, alpha, g, diag):
        return self.solver._lib.bgp_hodlr_grad_terms(self.solver._ptr, which, r, alpha, g, diag)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_computed"] = False
        state["_handle"] = None
        state.pop("solver", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

# -*- coding: utf-8 -*-
"""Pins the HODLR restatement (oracle/hodlr_oracle.cpp).  The reference's _hodlr cannot be built here (Eigen absent),
so the pins are the ones the reference's own tests use — dense linear algebra on the explicitly built matrix
(tests/test_solvers.py:45-55) — plus the docs' golden log-likelihood, the libstdc++ RNG words and the SURVEY App. B
pivot sequence.  'Parity unpinned' at the granularity of pivots for anything else."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def test_rng_stream_is_libstdcxx(oracle):
    assert list(oracle.mt19937_words(42, 4)) == [1608637542, 3421126067, 4083286876, 787846414]
    assert list(oracle.uniform_ints(42, [131072, 131071, 131070, 1000, 7, 1])) == [49091, 104403, 124610, 183, 5, 0]


def test_reference_solver_case(oracle):
    """tests/test_solvers.py:29-62 at tol=1e-10 + SURVEY.md App. B."""
    from george_b200 import kernels
    from george_b200._spec import flatten
    x = GOLD["solver300__x"]
    N = len(x)
    yerr = np.ones(N)
    kernel = 1.0 * kernels.ExpSquaredKernel(1.0)
    spec = flatten(kernel)
    h = oracle.HODLR(spec, x, yerr, min_size=100, tol=1e-10, seed=42, rng_mode=1)
    K = oracle.value_symmetric(spec, x[:, None]) + np.eye(N)
    sgn, ld = np.linalg.slogdet(K)
    assert sgn == 1.0 and np.allclose(h.log_determinant, ld)
    assert abs(h.log_determinant - 69.730382271778) < 1e-9
    # the reference's BasicSolver (LAPACK) on the same inputs, from the golden file
    assert abs(h.log_determinant - float(GOLD["solver300__logdet"])) < 1e-9
    y = np.sin(x)
    assert np.allclose(h.apply_inverse(y), GOLD["solver300__alpha"])
    assert np.allclose(h.dot_solve(y), float(GOLD["solver300__dot"]))
    assert np.allclose(h.apply_inverse(K), np.eye(N))
    nodes = h.nodes()
    assert [(n["start"], n["size"], n["is_leaf"]) for n in nodes] == [(0, 300, 0), (0, 150, 1), (150, 150, 1)]
    assert nodes[0]["rank"] == 14 and nodes[0]["rng_draws"] == 33
    rows, cols = h.pivots(0, 14)
    assert list(rows) == [56, 26, 86, 22, 62, 21, 13, 8, 45, 19, 2, 7, 0, 35]
    assert list(cols) == [149, 148, 144, 135, 146, 131, 124, 140, 118, 133, 112, 145, 106, 127]


def test_docs_golden_loglikelihood(oracle):
    """docs/tutorials/scaling.rst:56-91 -> 133.946394912 (N=100 < 2*min_size: a single leaf, exact)."""
    from george_b200 import kernels
    from george_b200._spec import flatten
    np.random.seed(1234)
    x = np.sort(np.random.uniform(0, 10, 50000))
    y = np.sin(x)
    k = np.var(y) * kernels.ExpSquaredKernel(1.0)
    n = 100
    sigma = np.sqrt(0.1 ** 2 + 1.25e-12) * np.ones(n)
    h = oracle.HODLR(flatten(k), x[:n], sigma)
    ll = -0.5 * (n * np.log(2 * np.pi) + h.log_determinant) - 0.5 * h.dot_solve(y[:n])
    assert abs(ll - 133.946394912) < 1e-8
    assert abs(ll - float(GOLD["docs__loglike_n100"])) < 1e-9


@pytest.mark.parametrize("n,min_size", [(777, 50), (2000, 100)])
def test_against_dense(oracle, n, min_size):
    from george_b200 import kernels
    from george_b200._spec import flatten
    rng = np.random.default_rng(n)
    x = np.sort(rng.uniform(0, 10 * n / 1000, n))
    yerr = 0.1 * np.ones(n)
    y = np.sin(x)
    for kernel in (1.0 * kernels.ExpSquaredKernel(1.0),
                   1.0 * kernels.ExpSquaredKernel(1.0) + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0))):
        spec = flatten(kernel)
        K = oracle.value_symmetric(spec, x[:, None]) + np.diag(yerr ** 2)
        for mode in (0, 1):
            h = oracle.HODLR(spec, x, yerr, min_size=min_size, tol=1e-10, rng_mode=mode)
            ld = np.linalg.slogdet(K)[1]
            assert abs(h.log_determinant - ld) <= 1e-9 * abs(ld)
            a = np.linalg.solve(K, y)
            assert np.linalg.norm(h.apply_inverse(y) - a) <= 1e-5 * np.linalg.norm(a)  # tol * cond(K)


def test_tree_geometry_rule(oracle):
    """hodlr.h:48-49: split iff size // 2 >= min_size; children (start, half) and (start+half, size-half)."""
    from george_b200 import kernels
    from george_b200._spec import flatten
    from george_b200.parallel import shard_ranges
    spec = flatten(1.0 * kernels.ExpSquaredKernel(1.0))
    for n, ms in [(199, 100), (200, 100), (201, 100), (1000, 100), (1023, 64)]:
        x = np.linspace(0, 10, n)
        nodes = oracle.HODLR(spec, x, np.ones(n), min_size=ms, tol=1e-3).nodes()
        for nd in nodes:
            assert nd["is_leaf"] == int(nd["size"] // 2 < ms)
        for p in (1, 2, 4):
            rr = shard_ranges(n, p, ms)
            cut = [(nd["start"], nd["size"]) for nd in nodes if nd["depth"] == p.bit_length() - 1]
            if rr is None:
                assert len(cut) != p or any(nd["is_leaf"] for nd in nodes if nd["depth"] < p.bit_length() - 1)
            else:
                assert rr == cut


def test_lowrank_exhaust_and_pernode_modes_of_the_oracle(oracle):
    """The two modes the CUDA path adds to the reference algorithm exist in the oracle too (so that every mode the device
    runs has a CPU twin).  exhaust = lowrank must agree with the reference's dense fallback to rounding — every row was
    verified below 1e-14 — while keeping ranks <= 3 for the exactly-rank-2 Matern-3/2; both must agree with dense LAPACK."""
    from george_b200 import kernels
    from george_b200._spec import flatten
    rng = np.random.default_rng(5)
    n = 1800
    x = np.sort(rng.uniform(0, 10 * n / 1000, n))
    yerr = 0.1 * np.ones(n)
    y = np.sin(x) + 0.1 * rng.normal(size=n)
    spec = flatten(1.0 * kernels.Matern32Kernel(1.0))
    K = oracle.value_symmetric(spec, x[:, None]) + np.diag(yerr ** 2)
    ld = np.linalg.slogdet(K)[1]
    q = y @ np.linalg.solve(K, y)
    res = {}
    for rng_mode in (1, 0):
        for exhaust in (0, 1):
            h = oracle.HODLR(spec, x, yerr, min_size=100, tol=1e-10, seed=42, rng_mode=rng_mode, exhaust=exhaust)
            nodes = [nd for nd in h.nodes() if not nd["is_leaf"]]
            res[(rng_mode, exhaust)] = (h.log_determinant, h.dot_solve(y), max(nd["rank"] for nd in nodes),
                                        sum(nd["dense_fallback"] for nd in nodes))
            assert abs(h.log_determinant - ld) <= 1e-10 * abs(ld)
            assert abs(h.dot_solve(y) - q) <= 1e-9 * abs(q)
    for rng_mode in (1, 0):
        dense, low = res[(rng_mode, 0)], res[(rng_mode, 1)]
        assert dense[3] > 0 and low[3] == dense[3]       # the same nodes run out of rows in both modes ...
        assert dense[2] >= 100 and low[2] <= 3           # ... the reference stores them densely, lowrank keeps rank <= 3
        assert abs(dense[0] - low[0]) <= 1e-11 * abs(ld)

# -*- coding: utf-8 -*-
"""Matrix-free consumers of the covariance function (csrc/kmat_ops.cu) and the fused gradient terms, through the C ABI.

* ``bgp_kmat_matvec`` vs the oracle-built matrix times the vector (rtol 1e-12: same evaluations, different summation
  order), incl. ragged / empty shapes, several right-hand sides and the column-split path used for few test points.
* ``bgp_kmat_gradient_contract`` and ``bgp_{dense,hodlr}_grad_terms`` vs ``einsum`` over the oracle's gradient tensor.
* The size-independent property at BASELINE.json's full size: K (K^-1 y) == y at N = 2^18 for the headline workload,
  with K applied matrix-free (the matrix itself would be 550 GB).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _zoo():
    from conftest import make_kernels
    return make_kernels()


def _points(kernel, n, rng):
    nd = kernel.ndim
    x = rng.uniform(-1.0, 1.0, (n, nd))
    return x[np.argsort(x[:, 0])]


@pytest.mark.parametrize("idx", range(14))
def test_matvec_matches_oracle_matrix(gpu, oracle, idx):
    from george_b200._spec import flatten
    name, kernel = _zoo()[idx]
    rng = np.random.default_rng(100 + idx)
    x1, x2 = _points(kernel, 77, rng), _points(kernel, 1301, rng)
    spec = flatten(kernel)
    K = oracle.value_general(spec, x1, x2)
    v = rng.normal(size=(1301, 5))
    out = kernel.matvec(x1, x2, v)
    assert out.shape == (77, 5)
    scale = np.abs(K) @ np.abs(v)
    assert np.all(np.abs(out - K @ v) <= 1e-12 * scale + 1e-300), name
    out1 = kernel.matvec(x1, x2, v[:, 0])
    assert out1.shape == (77,)
    assert np.all(np.abs(out1 - K @ v[:, 0]) <= 1e-12 * scale[:, 0] + 1e-300), name


@pytest.mark.parametrize("n1,n2", [(1, 1), (3, 5000), (64, 512), (65, 513), (700, 3), (1, 40000)])
def test_matvec_shapes_and_diag(gpu, oracle, n1, n2):
    from george_b200 import kernels as K
    from george_b200._spec import flatten
    rng = np.random.default_rng(n1 * 7 + n2)
    kernel = 1.3 * K.Matern32Kernel(0.7)
    x1 = np.sort(rng.uniform(0, 5, n1))[:, None]
    x2 = np.sort(rng.uniform(0, 5, n2))[:, None]
    Km = oracle.value_general(flatten(kernel), x1, x2)
    v = rng.normal(size=n2)
    out = kernel.matvec(x1, x2, v)
    assert np.all(np.abs(out - Km @ v) <= 1e-12 * (np.abs(Km) @ np.abs(v)))
    if n2 > 5000:   # the square check below builds an n2 x n2 matrix on the CPU
        return
    # square operator with a diagonal term: (K + diag) v
    d = rng.uniform(0.1, 1.0, n2)
    Ks = oracle.value_symmetric(flatten(kernel), x2)
    w = kernel.matvec(x2, x2, v, diag=d)
    assert np.all(np.abs(w - (Ks @ v + d * v)) <= 1e-12 * (np.abs(Ks) @ np.abs(v) + d * np.abs(v)))


def test_predict_mean_uses_matvec_and_matches_matrix_path(gpu):
    import george_b200 as george
    from george_b200 import kernels
    rng = np.random.default_rng(8)
    x = np.sort(rng.uniform(0, 20, 3000))
    y = np.sin(x) + 0.1 * rng.normal(size=x.size)
    t = np.linspace(-1, 21, 137)
    for solver, kw in ((george.BasicSolver, {}), (george.HODLRSolver, {"tol": 1e-12})):
        gp = george.GP(1.0 * kernels.ExpSquaredKernel(1.0), solver=solver, **kw)
        gp.compute(x, 0.1)
        mu = gp.predict(y, t, return_cov=False)
        mu_v, var = gp.predict(y, t, return_var=True)
        assert mu.shape == (137,)
        np.testing.assert_allclose(mu, mu_v, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("idx", range(14))
def test_gradient_contract_matches_einsum(gpu, oracle, idx):
    from george_b200._spec import flatten
    name, kernel = _zoo()[idx]
    rng = np.random.default_rng(200 + idx)
    n = 150
    x = _points(kernel, n, rng)
    A = rng.normal(size=(n, n))  # NOT symmetric: the contraction must use A_ij + A_ji
    mask = np.ones(len(kernel.get_parameter_vector(include_frozen=True)), dtype=bool)
    if mask.size > 1:
        mask[1] = False
    which = mask.astype(np.uint32)
    dK = oracle.gradient_general(flatten(kernel), which, x, x)
    ref = np.einsum("ijk,ij", dK, A)
    got = kernel.kernel.gradient_contract(which, x, A)
    scale = np.einsum("ijk,ij", np.abs(dK), np.abs(A))
    assert np.all(np.abs(got - ref) <= 1e-11 * scale + 1e-300), (name, got, ref)
    assert np.all(got[~mask] == 0.0)


@pytest.mark.parametrize("solver_name", ["basic", "hodlr"])
def test_grad_terms_match_host_composition(gpu, oracle, solver_name):
    """solver.grad_terms == (K^-1 r, einsum(dK, alpha alpha^T - K^-1), diag(...)) built from dense linear algebra."""
    import george_b200 as george
    from george_b200 import kernels
    from george_b200._spec import flatten
    rng = np.random.default_rng(31)
    n = 700
    x = rng.uniform(0, 1, (n, 2))
    x = x[np.argsort(x[:, 0])]
    kernel = 0.7 * kernels.Matern52Kernel([0.3, 0.6], ndim=2) + 0.2 * kernels.ExpSquaredKernel(0.1, ndim=2, axes=0)
    yerr = 0.1 * np.ones(n)
    r = np.sin(4 * x[:, 0]) + x[:, 1]
    spec = flatten(kernel)
    Kd = oracle.value_symmetric(spec, x) + np.diag(yerr ** 2)
    Kinv = np.linalg.inv(Kd)
    alpha = Kinv @ r
    A = np.outer(alpha, alpha) - Kinv
    np_ = len(kernel.get_parameter_vector(include_frozen=True))
    which = np.ones(np_, dtype=np.uint32)
    dK = oracle.gradient_general(spec, which, x, x)
    ref = np.einsum("ijk,ij", dK, A)
    if solver_name == "basic":
        s = george.BasicSolver(kernel)
    else:
        s = george.HODLRSolver(kernel, tol=1e-12)
    s.compute(x, yerr)
    a, g, dA = s.grad_terms(r, which)
    assert np.linalg.norm(a - alpha) <= 1e-8 * np.linalg.norm(alpha)
    scale = np.einsum("ijk,ij", np.abs(dK), np.abs(A))
    assert np.all(np.abs(g - ref) <= 1e-7 * scale), (g, ref)
    assert np.linalg.norm(dA - np.diag(A)) <= 1e-7 * np.linalg.norm(np.diag(A))


def test_dense_few_rhs_solve_large(gpu, oracle):
    """The few-right-hand-side substitution (csrc/dense.cu: trsv_*_step_kernel) at a size with many blocks, ragged."""
    import george_b200 as george
    from george_b200 import kernels
    from george_b200._spec import flatten
    rng = np.random.default_rng(77)
    n = 2500 + 37
    x = np.sort(rng.uniform(0, 30, n))[:, None]
    kernel = 1.0 * kernels.Matern32Kernel(2.0)
    yerr = 0.2 * np.ones(n)
    Kd = oracle.value_symmetric(flatten(kernel), x) + np.diag(yerr ** 2)
    s = george.BasicSolver(kernel)
    s.compute(x, yerr)
    for k in (1, 2, 4, 7, 8, 9):
        Y = rng.normal(size=(n, k))
        got = s.apply_inverse(Y if k > 1 else Y[:, 0])
        ref = np.linalg.solve(Kd, Y)
        assert np.linalg.norm(got.reshape(n, k) - ref) <= 1e-9 * np.linalg.norm(ref), k
    y = rng.normal(size=n)
    assert abs(s.dot_solve(y) - y @ np.linalg.solve(Kd, y)) <= 1e-9 * abs(y @ np.linalg.solve(Kd, y))


def test_full_size_round_trip(gpu):
    """BASELINE.json's metric configuration (Matern32 1-D, N = 2^18, leaf 256, tol 1e-10): b = K^-1 y from the HODLR
    factorisation, then K b with K = k(x, x) + diag applied matrix-free, must give y back.  Also checks the two RNG-free
    invariants of the factorisation: log-det is finite and dot_solve(y) == y . b."""
    import george_b200 as george
    from george_b200 import kernels
    n = 262144
    rng = np.random.default_rng(1234)
    x = np.sort(rng.uniform(0, 10 * n / 1000, n))
    yerr = 0.1 * np.ones(n)
    y = np.sin(x) + 0.1 * rng.normal(size=n)
    kernel = 1.0 * kernels.Matern32Kernel(1.0)
    s = george.HODLRSolver(kernel, min_size=256, tol=1e-10, seed=42, exhaust="lowrank")
    s.compute(x[:, None], yerr)
    assert np.isfinite(s.log_determinant)
    b = s.apply_inverse(y)[:, 0]
    back = kernel.matvec(x[:, None], x[:, None], b, diag=yerr ** 2)
    assert np.linalg.norm(back - y) <= 1e-8 * np.linalg.norm(y)
    assert abs(s.dot_solve(y) - y @ b) <= 1e-10 * abs(y @ b)

# -*- coding: utf-8 -*-
"""
Generates tests/golden/*.npz by IMPORTING THE REFERENCE in the build container (it cannot travel to the GPU box).

Run from the repo root:   python tests/golden/make_golden.py

What is used from /root/reference: its unmodified Python package (GP, kernels, BasicSolver, scipy LAPACK path) with its
own kernel_interface.cpp compiled by oracle/Makefile (`make -C oracle ref`).  The `_hodlr` extension cannot be built
(Eigen is an absent submodule), so a stub module is injected; HODLR goldens therefore come from dense linear algebra on
reference-built matrices (what the reference's own tests compare against, tests/test_solvers.py:45-55).
The staging directory lives under /tmp: no reference source is copied into the repository.
"""
import glob
import os
import shutil
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/src/george"
STAGE = "/tmp/george_ref_stage"


def stage_reference():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "kernel_interface*.so"))
    assert so, "run `make -C oracle ref` first"
    if os.path.exists(STAGE):
        shutil.rmtree(STAGE)
    pkg = os.path.join(STAGE, "george")
    os.makedirs(os.path.join(pkg, "solvers"))
    for f in ("__init__.py", "gp.py", "kernels.py", "metrics.py", "modeling.py", "utils.py"):
        os.symlink(os.path.join(REF, f), os.path.join(pkg, f))
    for f in ("__init__.py", "basic.py", "hodlr.py", "trivial.py"):
        os.symlink(os.path.join(REF, "solvers", f), os.path.join(pkg, "solvers", f))
    os.symlink(so[0], os.path.join(pkg, os.path.basename(so[0])))
    with open(os.path.join(pkg, "george_version.py"), "w") as fh:
        fh.write("version = '0.0.0+ref'\n")
    with open(os.path.join(pkg, "solvers", "_hodlr.py"), "w") as fh:
        fh.write("class HODLRSolver(object):\n    def __init__(self):\n        raise RuntimeError('needs Eigen')\n")
    sys.path.insert(0, STAGE)
    import george
    return george


def main():
    george = stage_reference()
    from george import kernels
    out = {}
    rng = np.random.default_rng(20260923)

    # ---- kernel values + gradients of the hot kernels (value_general / gradient_general / value_symmetric) ----
    zoo = {
        "expsq_1d": (lambda: 1.0 * kernels.ExpSquaredKernel(1.0), 1),
        "m32_1d": (lambda: 2.3 * kernels.Matern32Kernel(0.7), 1),
        "m52_3d": (lambda: 1.0 * kernels.Matern52Kernel(0.5, ndim=3), 3),
        "cfg5_1d": (lambda: 1.0 * kernels.ExpSquaredKernel(1.0)
                    + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=np.log(3.0)), 1),
        "m52_3d_axis": (lambda: kernels.Matern52Kernel([0.5, 1.0, 2.0], ndim=3), 3),
        "expsq_3d_general": (lambda: kernels.ExpSquaredKernel([[1.0, 0.1, 0.2], [0.1, 2.0, 0.3], [0.2, 0.3, 1.5]],
                                                             ndim=3), 3),
        "ratquad": (lambda: kernels.RationalQuadraticKernel(log_alpha=0.3, metric=1.2, ndim=3), 3),
    }
    for name, (mk, nd) in zoo.items():
        k = mk()
        x1 = rng.normal(size=(23, nd))
        x2 = rng.normal(size=(17, nd))
        out[name + "__x1"] = x1
        out[name + "__x2"] = x2
        out[name + "__value"] = k.get_value(x1, x2)
        out[name + "__sym"] = k.get_value(x1)
        out[name + "__grad"] = k.get_gradient(x1, x2, include_frozen=True)

    # ---- BasicSolver path: reference tests/test_solvers.py:29-58 setup ----
    np.random.seed(1234)
    N = 300
    x = np.sort(10 * np.random.randn(N))
    yerr = np.ones(N)
    kernel = 1.0 * kernels.ExpSquaredKernel(1.0)
    s = george.BasicSolver(kernel)
    s.compute(x[:, None], yerr)
    y = np.sin(x)
    out["solver300__x"] = x
    out["solver300__logdet"] = np.array(s.log_determinant)
    out["solver300__alpha"] = s.apply_inverse(y)
    out["solver300__dot"] = np.array(s.dot_solve(y))

    # ---- docs golden: docs/tutorials/scaling.rst:56-91, log-likelihood 133.946394912 at N=100 ----
    np.random.seed(1234)
    xx = np.sort(np.random.uniform(0, 10, 50000))
    yy = np.sin(xx)
    k = np.var(yy) * kernels.ExpSquaredKernel(1.0)
    gp = george.GP(k)
    gp.compute(xx[:100], 0.1 * np.ones(100))
    out["docs__loglike_n100"] = np.array(gp.log_likelihood(yy[:100]))

    # ---- GP.predict through the reference (tests/test_gp.py:59-83 setup) ----
    np.random.seed(42)
    kern = kernels.ExpSquaredKernel(1.0)
    kern.freeze_all_parameters()
    gp = george.GP(kern, white_noise=0.0)
    x0 = np.linspace(-10, 10, 50)
    xs = np.sort(np.random.uniform(-10, 10, 300))
    gp.compute(xs)
    mu, cov = gp.predict(np.sin(xs), x0)
    out["predict__x"] = xs
    out["predict__x0"] = x0
    out["predict__mu"] = mu
    out["predict__cov"] = cov

    # ---- config 4 in miniature: Matern52 3-D BasicSolver log-likelihood ----
    r2 = np.random.default_rng(4)
    x3 = r2.uniform(0, 1, (400, 3))
    x3 = x3[np.argsort(x3[:, 0])]
    y3 = np.sin(x3.sum(axis=1))
    gp = george.GP(1.0 * kernels.Matern52Kernel(0.5, ndim=3))
    gp.compute(x3, 0.1)
    out["cfg4mini__x"] = x3
    out["cfg4mini__y"] = y3
    out["cfg4mini__loglike"] = np.array(gp.log_likelihood(y3))

    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays; docs loglike =", float(out["docs__loglike_n100"]))


if __name__ == "__main__":
    main()

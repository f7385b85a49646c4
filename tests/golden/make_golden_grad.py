# -*- coding: utf-8 -*-
"""
Generates tests/golden/reference_golden_grad.npz by IMPORTING THE REFERENCE in the build container: the gradient of the
log-likelihood, multi-RHS apply_inverse and mean-only predictions computed by the reference's own GP / BasicSolver
(src/george/gp.py:406-468, 277-301, 482-545) on the setup of its tests/test_gp.py:16-56.

Run from the repo root:   python tests/golden/make_golden_grad.py     (needs `make -C oracle ref`)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import ROOT, stage_reference  # noqa: E402


def main():
    george = stage_reference()
    from george import kernels
    out = {}
    np.random.seed(123)
    N, ndim = 305, 3
    x = np.random.rand(N, ndim)
    x = x[np.argsort(x[:, 0])]
    y = np.sin(np.sum(x, axis=1))
    out["x"], out["y"] = x, y
    cases = {
        "plain": dict(),
        "white": dict(white_noise=0.1, fit_white_noise=True),
        "white_mean": dict(white_noise=-2.0, fit_white_noise=True, mean=0.3, fit_mean=True),
    }
    for name, kw in cases.items():
        kernel = 0.1 * kernels.ExpSquaredKernel(0.5, ndim=ndim)
        gp = george.GP(kernel, **kw)
        gp.compute(x, yerr=0.1)
        out[name + "__names"] = np.array(gp.get_parameter_names())
        out[name + "__vector"] = gp.get_parameter_vector()
        out[name + "__loglike"] = np.array(gp.log_likelihood(y))
        out[name + "__grad"] = gp.grad_log_likelihood(y)
    # a sum kernel with a frozen parameter and an axis-aligned metric (5 kernel parameters, 4 active)
    kernel = 0.5 * kernels.Matern32Kernel([0.3, 0.6, 1.2], ndim=ndim) + 0.05 * kernels.ExpSquaredKernel(0.2, ndim=ndim, axes=0)
    kernel.freeze_parameter("k2:k1:log_constant")
    gp = george.GP(kernel)
    gp.compute(x, yerr=0.05)
    out["sum__names"] = np.array(gp.get_parameter_names())
    out["sum__loglike"] = np.array(gp.log_likelihood(y))
    out["sum__grad"] = gp.grad_log_likelihood(y)
    Y = np.vstack([y, np.cos(3 * y), y ** 2]).T
    out["sum__apply_inverse3"] = gp.apply_inverse(Y)
    t = np.random.rand(40, ndim)
    out["sum__t"] = t
    out["sum__mu"] = gp.predict(y, t, return_cov=False)
    mu, var = gp.predict(y, t, return_var=True)
    out["sum__var"] = var
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_grad.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()

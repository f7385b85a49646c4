# -*- coding: utf-8 -*-
"""
The ``GP`` object: caller of the accelerated path.

Behavioural mirror of the reference's ``src/george/gp.py:22-643`` (same constructor, methods, caching and error
semantics) — ``compute`` hands ``(x, sqrt(yerr^2 + exp(white_noise)))`` to a freshly built solver plugin
(``gp.py:303-337``), ``log_likelihood`` is ``_const - 0.5 * solver.dot_solve(y - mean)`` (``gp.py:369-397``) and
``predict`` reuses the factorisation (``gp.py:482-545``).  All O(N^2)/O(N log^2 N) work happens inside the solver
plugins on the B200; this file is host bookkeeping.
"""

import warnings

import numpy as np
from numpy.linalg import LinAlgError

from . import kernels
from .modeling import ConstantModel, ModelSet
from .solvers import BasicSolver, TrivialSolver
from .utils import multivariate_gaussian_samples

__all__ = ["GP"]

# jitter added to the diagonal when no observational uncertainty is given (reference gp.py:19)
TINY = 1.25e-12


def _as_model(obj):
    """Scalars become ConstantModel; anything else is assumed to follow the modeling protocol."""
    try:
        value = float(obj)
    except TypeError:
        return obj
    return ConstantModel(float(value))


def _is_number(obj):
    try:
        float(obj)
    except TypeError:
        return False
    return True


class GP(ModelSet):
    """Gaussian-process regression model.

    :param kernel: a :class:`kernels.Kernel` (default: ``EmptyKernel``)
    :param fit_kernel: include the kernel parameters in the parameter vector (default ``True``)
    :param mean: scalar, callable-model or modeling-protocol object (default ``0``)
    :param fit_mean: fit the mean parameters (default: only when a non-scalar mean is given)
    :param white_noise: log of the white-noise variance added on the diagonal (default ``log(TINY)``)
    :param fit_white_noise: fit the white-noise parameters
    :param solver: solver plugin class (default ``BasicSolver``, or ``TrivialSolver`` without a kernel)
    :param kwargs: forwarded to the solver constructor (e.g. ``min_size``, ``tol``, ``seed`` for ``HODLRSolver``)
    """

    def __init__(self, kernel=None, fit_kernel=True, mean=None, fit_mean=None, white_noise=None,
                 fit_white_noise=None, solver=None, **kwargs):
        self._computed = False
        self._alpha = None
        self._y = None

        super(GP, self).__init__([
            ("mean", ConstantModel(0.0) if mean is None else _as_model(mean)),
            ("white_noise", ConstantModel(np.log(TINY)) if white_noise is None else _as_model(white_noise)),
            ("kernel", kernels.EmptyKernel() if kernel is None else kernel),
        ])

        # a plain number for mean / white_noise is not fitted unless asked for
        if _is_number(mean) and fit_mean is None:
            fit_mean = False
        if _is_number(white_noise) and fit_white_noise is None:
            fit_white_noise = False

        if not fit_kernel:
            self.models["kernel"].freeze_all_parameters()
        if mean is None or (fit_mean is not None and not fit_mean):
            self.models["mean"].freeze_all_parameters()
        if white_noise is None or (fit_white_noise is not None and not fit_white_noise):
            self.models["white_noise"].freeze_all_parameters()

        if solver is None:
            no_kernel = kernel is None or kernel.kernel_type == kernels.EmptyKernel.kernel_type
            solver = TrivialSolver if no_kernel else BasicSolver
        self.solver_type = solver
        self.solver_kwargs = kwargs
        self.solver = None

    # -- sub-models ------------------------------------------------------------------------------------------------
    @property
    def mean(self):
        return self.models["mean"]

    @property
    def white_noise(self):
        return self.models["white_noise"]

    @staticmethod
    def _model_input(x):
        return x[:, 0] if (x.ndim == 2 and x.shape[1] == 1) else x

    def _call_mean(self, x):
        mu = self.mean.get_value(self._model_input(x)).flatten()
        if not np.all(np.isfinite(mu)):
            raise ValueError("mean function returned NaN or Inf for parameters:\n{0}".format(
                self.mean.get_parameter_dict(include_frozen=True)))
        return mu

    def _call_mean_gradient(self, x):
        g = self.mean.get_gradient(self._model_input(x))
        if np.any(np.isnan(g)) or np.any(np.isinf(g)):
            raise ValueError("mean gradient function returned NaN or Inf for parameters:\n{0}".format(
                self.mean.get_parameter_dict(include_frozen=True)))
        return g

    def _call_white_noise(self, x):
        return self.white_noise.get_value(self._model_input(x)).flatten()

    # Constant mean / white-noise models (the defaults) need no per-point array: at N ~ 10^5..10^6 every extra pass
    # over an N-vector on the host costs about as much as a level of the factorisation on the device.
    def _constant_of(self, model):
        return float(model.value) if type(model) is ConstantModel else None

    def _sigma(self, x):
        """sqrt(yerr^2 + exp(white_noise(x)))  (reference gp.py:330)."""
        c = self._constant_of(self.white_noise)
        if c is None:
            return np.sqrt(self._yerr2 + np.exp(self._call_white_noise(x)))
        sigma = self._yerr2 + np.exp(c)
        return np.sqrt(sigma, out=sigma)

    def _residual_of(self, y):
        """y - mean(x) as a contiguous float64 vector (reference gp.py:388-393); raises on a non-finite mean."""
        c = self._constant_of(self.mean)
        if c is None:
            return np.ascontiguousarray(self._check_dimensions(y) - self._call_mean(self._x), dtype=np.float64)
        if not np.isfinite(c):
            raise ValueError("mean function returned NaN or Inf for parameters:\n{0}".format(
                self.mean.get_parameter_dict(include_frozen=True)))
        y = self._check_dimensions(y)
        if c == 0.0:
            return np.ascontiguousarray(y, dtype=np.float64)  # (no copy when y already is one)
        return np.ascontiguousarray(y - c, dtype=np.float64)

    def _call_white_noise_gradient(self, x):
        return self.white_noise.get_gradient(self._model_input(x))

    # -- state -----------------------------------------------------------------------------------------------------
    @property
    def computed(self):
        """Is the factorisation current w.r.t. the kernel parameters?"""
        return self._computed and self.solver.computed and (self.kernel is None or not self.kernel.dirty)

    @computed.setter
    def computed(self, v):
        self._computed = v
        if v and self.kernel is not None:
            self.kernel.dirty = False

    def parse_samples(self, t):
        """Coerce coordinates to ``(nsamples, ndim)``; 1-D input means one-dimensional samples."""
        t = np.atleast_1d(t)
        if t.ndim == 1:
            t = np.atleast_2d(t).T
        if t.ndim != 2 or (self.kernel is not None and t.shape[1] != self.kernel.ndim):
            raise ValueError("Dimension mismatch")
        return t

    def _check_dimensions(self, y, check_dim=True):
        n = self._x.shape[0]
        y = np.atleast_1d(y)
        if check_dim and y.ndim > 1:
            raise ValueError("The predicted dimension must be 1-D")
        if len(y) != n:
            raise ValueError("Dimension mismatch")
        return y

    def _residual(self, y):
        return np.ascontiguousarray(self._check_dimensions(y) - self._call_mean(self._x), dtype=np.float64)

    def _compute_alpha(self, y, cache):
        """alpha = K^-1 (y - mean); cached on the identity of y's values (gp.py:260-275)."""
        if not cache:
            return self.solver.apply_inverse(self._residual(y), in_place=True).flatten()
        if self._alpha is None or not np.array_equiv(y, self._y):
            self._y = y
            self._alpha = self.solver.apply_inverse(self._residual(y), in_place=True).flatten()
        return self._alpha

    def apply_inverse(self, y):
        """``K^-1 (y - mean)`` for a vector or an ``(nsamples, K)`` matrix."""
        self.recompute(quiet=False)
        r = np.array(y, dtype=np.float64, order="F")
        r = self._check_dimensions(r, check_dim=False)
        mu = self._call_mean(self._x)
        r -= mu.reshape((-1,) + (1,) * (r.ndim - 1))
        b = self.solver.apply_inverse(r, in_place=True)
        return b.flatten() if r.ndim == 1 else b

    # -- the hot path ------------------------------------------------------------------------------------------------
    def compute(self, x, yerr=0.0, **kwargs):
        """Build and factorise the covariance matrix for coordinates ``x`` and uncertainties ``yerr``."""
        self._x = np.ascontiguousarray(self.parse_samples(x), dtype=np.float64)
        try:
            self._yerr2 = float(yerr) ** 2 * np.ones(len(x))
        except TypeError:
            self._yerr2 = self._check_dimensions(yerr) ** 2
        self._yerr2 = np.ascontiguousarray(self._yerr2, dtype=np.float64)

        # a new solver per compute: the factorisation is a snapshot of the current parameters
        self.solver = self.solver_type(self.kernel, **(self.solver_kwargs))
        self.solver.compute(self._x, self._sigma(self._x), **kwargs)

        self._const = -0.5 * (len(self._x) * np.log(2 * np.pi) + self.solver.log_determinant)
        self.computed = True
        self._alpha = None

    def recompute(self, quiet=False, **kwargs):
        """Refactorise if the kernel changed since the last ``compute``.  With ``quiet`` a failed factorisation
        returns ``False`` instead of raising."""
        if self.computed:
            return True
        if not (hasattr(self, "_x") and hasattr(self, "_yerr2")):
            raise RuntimeError("You need to compute the model first")
        try:
            self.compute(self._x, np.sqrt(self._yerr2), **kwargs)
        except (ValueError, LinAlgError):
            if quiet:
                return False
            raise
        return True

    def log_likelihood(self, y, quiet=False):
        """Marginal log-likelihood of ``y`` at the computed coordinates; ``-inf`` on failure when ``quiet``."""
        if not self.recompute(quiet=quiet):
            return -np.inf
        try:
            r = self._residual_of(y)
        except ValueError as exc:
            if quiet and "mean function" in str(exc):
                return -np.inf
            raise
        ll = self._const - 0.5 * self.solver.dot_solve(r)
        return ll if np.isfinite(ll) else -np.inf

    def lnlikelihood(self, y, quiet=False):
        warnings.warn("'lnlikelihood' is deprecated. Use 'log_likelihood'", DeprecationWarning)
        return self.log_likelihood(y, quiet=quiet)

    def grad_log_likelihood(self, y, quiet=False):
        """Gradient of :func:`log_likelihood` w.r.t. the active parameter vector (gp.py:406-468)."""
        nothing = np.zeros(len(self), dtype=np.float64)
        if not self.recompute(quiet=quiet):
            return nothing
        n_wn, n_k, n_mean = len(self.white_noise), len(self.kernel), len(self.mean)
        fused = getattr(self.solver, "grad_terms", None) if (n_wn or n_k) else None
        try:
            if fused is not None:
                # alpha, the kernel-gradient contraction and diag(alpha alpha^T - K^-1) in one device pass: neither
                # K^-1 nor the (N, N, P) gradient tensor visits the host (reference gp.py:437-466 forms both)
                mask = self.kernel.unfrozen_mask
                terms = fused(self._residual(y), mask.astype(np.uint32))
                if terms is None:
                    fused = None
                else:
                    alpha, gk, diagA = terms
            if fused is None:
                alpha = self._compute_alpha(y, False)
        except ValueError:
            if quiet:
                return nothing
            raise

        if fused is None and (n_wn or n_k):
            A = np.outer(alpha, alpha) - self.solver.get_inverse()
            diagA = np.diag(A)

        grad = np.empty(len(self))
        pos = 0
        if n_mean:
            try:
                dmu = self._call_mean_gradient(self._x)
            except ValueError:
                if quiet:
                    return nothing
                raise
            grad[pos:pos + n_mean] = np.dot(dmu, alpha)
            pos += n_mean
        if n_wn:
            wn = self._call_white_noise(self._x)
            dwn = self._call_white_noise_gradient(self._x)
            grad[pos:pos + n_wn] = 0.5 * np.sum((np.exp(wn) * diagA)[None, :] * dwn, axis=1)
            pos += n_wn
        if n_k:
            if fused is not None:
                grad[pos:pos + n_k] = 0.5 * gk[mask]
            else:
                # plug-in solvers without grad_terms: K^-1 comes from the solver, the contraction still runs on the
                # device without the (N, N, P) tensor
                mask = self.kernel.unfrozen_mask
                grad[pos:pos + n_k] = 0.5 * self.kernel.kernel.gradient_contract(mask.astype(np.uint32), self._x, A)[mask]
        return grad

    def grad_lnlikelihood(self, y, quiet=False):
        warnings.warn("'grad_lnlikelihood' is deprecated. Use 'grad_log_likelihood'", DeprecationWarning)
        return self.grad_log_likelihood(y, quiet=quiet)

    def nll(self, vector, y, quiet=True):
        self.set_parameter_vector(vector)
        if not np.isfinite(self.log_prior()):
            return np.inf
        return -self.log_likelihood(y, quiet=quiet)

    def grad_nll(self, vector, y, quiet=True):
        self.set_parameter_vector(vector)
        if not np.isfinite(self.log_prior()):
            return np.zeros(len(vector))
        return -self.grad_log_likelihood(y, quiet=quiet)

    def predict(self, y, t, return_cov=True, return_var=False, cache=True, kernel=None):
        """Conditional predictive distribution at ``t``: ``mu``, ``(mu, cov)`` or ``(mu, var)``."""
        self.recompute()
        alpha = self._compute_alpha(y, cache)
        xs = self.parse_samples(t)
        if kernel is None:
            kernel = self.kernel

        if not (return_var or return_cov):
            # mean only: K(x*, x) alpha evaluated matrix-free on the device (csrc/kmat_ops.cu); the reference forms the
            # (n*, N) matrix on the host (gp.py:524-528), which stops being possible long before N = 2^18
            return kernel.matvec(xs, self._x, alpha) + self._call_mean(xs)

        Kxs = kernel.get_value(xs, self._x)
        mu = np.dot(Kxs, alpha) + self._call_mean(xs)

        KinvKxs = self.solver.apply_inverse(Kxs.T)
        if return_var:
            var = kernel.get_value(xs, diag=True)
            var -= np.sum(Kxs.T * KinvKxs, axis=0)
            return mu, var
        cov = kernel.get_value(xs)
        cov -= np.dot(Kxs, KinvKxs)
        return mu, cov

    def sample_conditional(self, y, t, size=1):
        mu, cov = self.predict(y, t)
        return multivariate_gaussian_samples(cov, size, mean=mu)

    def sample(self, t=None, size=1):
        """Draw from the prior, at ``t`` or (``t is None``) at the computed coordinates via the Cholesky factor."""
        if t is None:
            self.recompute()
            n = self._x.shape[0]
            draws = self.solver.apply_sqrt(np.random.randn(size, n))
            draws += self._call_mean(self._x)
            return draws[0] if size == 1 else draws
        x = self.parse_samples(t)
        cov = self.get_matrix(x)
        cov[np.diag_indices_from(cov)] += TINY
        return multivariate_gaussian_samples(cov, size, mean=self._call_mean(x))

    def get_matrix(self, x1, x2=None):
        x1 = self.parse_samples(x1)
        if x2 is None:
            return self.kernel.get_value(x1)
        return self.kernel.get_value(x1, self.parse_samples(x2))

    # modeling-protocol synonyms
    def get_value(self, *args, **kwargs):
        return self.log_likelihood(*args, **kwargs)

    def get_gradient(self, *args, **kwargs):
        return self.grad_log_likelihood(*args, **kwargs)

# -*- coding: utf-8 -*-
"""
Distance metrics of the stationary kernels and the axis sub-space of the others.

Same public surface and parameterisation as the reference's ``src/george/metrics.py:14-140``:

* scalar ``metric``      -> isotropic,   ``metric_type = 0``, one parameter ``log_M_0_0 = log(metric)``
* 1-D ``metric``         -> axis aligned, ``metric_type = 1``, ``log_M_i_i = log(metric[i])``
* 2-D ``metric`` (SPD)   -> general,      ``metric_type = 2``, packed lower Cholesky factor with the diagonal
                             stored as logs (row-major lower-triangle order, the layout
                             ``metrics.h:166-180`` expects)

``r^2 = (x1-x2)^T M^{-1} (x1-x2)`` restricted to ``axes``.  The numeric work happens on the device
(``csrc/kernel_eval.cuh``); this class only carries parameters.
"""

import numpy as np
from scipy.linalg import cho_factor

from .modeling import Model

__all__ = ["Metric", "Subspace"]


class Subspace(object):
    """``ndim`` input dimensions of which only ``axes`` are used."""

    def __init__(self, ndim, axes=None):
        self.ndim = int(ndim)
        self.axes = np.atleast_1d(np.arange(self.ndim) if axes is None else axes).astype(int)
        if np.any(self.axes >= self.ndim):
            raise ValueError("invalid axis for {0} dimensional metric".format(self.ndim))


class Metric(Model):

    def __init__(self, metric, bounds=None, ndim=None, axes=None, lower=True):
        if isinstance(metric, Metric):  # copy constructor
            self.metric_type = metric.metric_type
            self.parameter_names = metric.parameter_names
            self.unfrozen_mask = metric.unfrozen_mask
            self.set_parameter_vector(metric.get_parameter_vector(include_frozen=True), include_frozen=True)
            self.parameter_bounds = metric.parameter_bounds
            self.ndim = metric.ndim
            self.axes = metric.axes
            return

        if ndim is None:
            raise ValueError("missing required parameter 'ndim'")
        sub = Subspace(ndim, axes=axes)
        self.ndim, self.axes = sub.ndim, sub.axes
        naxes = len(self.axes)

        names, values = [], []
        try:
            scalar = float(metric)
        except TypeError:
            m = np.atleast_1d(metric)
            if m.ndim == 1:
                self.metric_type = 1
                if len(m) != naxes:
                    raise ValueError("dimension mismatch")
                if np.any(m <= 0.0):
                    raise ValueError("invalid (negative) metric")
                for i, v in enumerate(m):
                    names.append("log_M_{0}_{0}".format(i))
                    values.append(np.log(v))
            elif m.ndim == 2:
                self.metric_type = 2
                if m.shape[0] != m.shape[1]:
                    raise ValueError("metric must be square")
                if len(m) != naxes:
                    raise ValueError("dimension mismatch")
                chol = cho_factor(m, lower=True)[0]
                d = np.diag_indices_from(chol)
                chol[d] = np.log(chol[d])
                packed = chol[np.tril_indices_from(chol)]
                k = 0
                for i in range(naxes):
                    names.append("log_L_{0}_{0}".format(i))
                    values.append(packed[k])
                    k += 1
                    for j in range(i + 1, naxes):
                        names.append("L_{0}_{1}".format(i, j))
                        values.append(packed[k])
                        k += 1
            else:
                raise ValueError("invalid metric dimensions")
        else:
            self.metric_type = 0
            names.append("log_M_0_0")
            values.append(np.log(scalar))

        self.parameter_names = tuple(names)
        kwargs = dict(zip(names, values))
        if bounds is not None:
            kwargs["bounds"] = bounds
        super(Metric, self).__init__(**kwargs)

    def to_matrix(self):
        v = self.get_parameter_vector(include_frozen=True)
        naxes = len(self.axes)
        if self.metric_type == 0:
            return np.exp(v) * np.eye(naxes)
        if self.metric_type == 1:
            return np.diag(np.exp(v))
        L = np.zeros((naxes, naxes))
        L[np.tril_indices_from(L)] = v
        d = np.diag_indices_from(L)
        L[d] = np.exp(L[d])
        return np.dot(L, L.T)

    def __repr__(self):
        v = self.get_parameter_vector(include_frozen=True)
        if self.metric_type == 0:
            head = "{0}".format(float(np.exp(v[0])))
        elif self.metric_type == 1:
            head = "{0}".format(repr(np.exp(v)))
        else:
            head = "{0}".format(repr(self.to_matrix().tolist()))
        bounds = [(None if lo is None else np.exp(lo), None if hi is None else np.exp(hi))
                  for lo, hi in self.get_parameter_bounds(include_frozen=True)]
        return "Metric({0}, ndim={1}, axes={2}, bounds={3})".format(head, self.ndim, repr(self.axes), bounds)

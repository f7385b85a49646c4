# -*- coding: utf-8 -*-
"""
The parameter-vector ("modeling") protocol that GP, kernels, metrics and mean functions share.

Host-side bookkeeping only — no array work happens here.  It mirrors the public behaviour of the
reference's ``src/george/modeling.py`` (``Model`` :11-343, ``ModelSet`` :346-473, ``ConstantModel`` :476-491,
``CallableModel`` :494-507): named parameters stored as attributes, a boolean thaw mask, optional
``(min, max)`` bounds that define a flat ``log_prior``, a ``dirty`` flag that ``GP`` uses to decide when the
device factorisation must be rebuilt, and ``name:sub`` addressing inside composite models.
"""

from collections import OrderedDict

import numpy as np

__all__ = ["Model", "ModelSet", "ConstantModel", "CallableModel"]

_FD_STEP = 1.254e-5  # forward-difference step of the default gradient (reference modeling.py:120)


class Model(object):
    """A bag of named scalar parameters.

    Subclasses list their parameters in ``parameter_names``; values are given positionally or by keyword::

        class Line(Model):
            parameter_names = ("slope", "intercept")
            def get_value(self, x):
                return self.slope * x + self.intercept

    ``bounds`` may be a dict ``{name: (lo, hi)}`` or a list with one ``(lo, hi)`` pair per parameter.
    """

    parameter_names = tuple()

    def __init__(self, *args, **kwargs):
        nfull = self.full_size
        self.unfrozen_mask = np.ones(nfull, dtype=bool)
        self.dirty = True

        bounds = kwargs.pop("bounds", None)
        if bounds is None:
            bounds = {}
        if hasattr(bounds, "get"):
            self.parameter_bounds = [bounds.get(k, (None, None)) for k in self.parameter_names]
        else:
            self.parameter_bounds = list(bounds)
        if len(self.parameter_bounds) != nfull:
            raise ValueError("the number of bounds must equal the number of parameters")
        for b in self.parameter_bounds:
            if len(b) != 2:
                raise ValueError("the bounds for each parameter must have the format: '(min, max)'")

        if args:
            if len(args) != nfull:
                raise ValueError("expected {0} arguments but got {1}".format(nfull, len(args)))
            if kwargs:
                raise ValueError("parameters must be fully specified by arguments or keyword arguments, not both")
            self.parameter_vector = args
        else:
            values = []
            for k in self.parameter_names:
                v = kwargs.pop(k, None)
                if v is None:
                    raise ValueError("missing parameter '{0}'".format(k))
                values.append(v)
            self.parameter_vector = values
            if kwargs:
                raise ValueError("unrecognized parameter(s) '{0}'".format(list(kwargs.keys())))

        if not np.isfinite(self.log_prior()):
            raise ValueError("non-finite log prior value")

    # -- value / gradient -------------------------------------------------------------------------------------
    def get_value(self, *args, **kwargs):
        raise NotImplementedError("overloaded by subclasses")

    def compute_gradient(self, *args, **kwargs):
        """First-order forward differences; subclasses overload with something analytic."""
        # over the FULL vector so that get_gradient's thaw mask lines up (the reference differentiates only the thawed
        # entries and then masks again, modeling.py:121-137, which breaks as soon as a parameter is frozen)
        p = self.get_parameter_vector(include_frozen=True)
        f0 = self.get_value(*args, **kwargs)
        g = np.empty([len(p)] + list(np.shape(f0)), dtype=np.float64)
        for i, pi in enumerate(p):
            p[i] = pi + _FD_STEP
            self.set_parameter_vector(p, include_frozen=True)
            g[i] = (self.get_value(*args, **kwargs) - f0) / _FD_STEP
            p[i] = pi
            self.set_parameter_vector(p, include_frozen=True)
        return g

    def get_gradient(self, *args, **kwargs):
        include_frozen = kwargs.pop("include_frozen", False)
        g = self.compute_gradient(*args, **kwargs)
        return g if include_frozen else g[self.unfrozen_mask]

    # -- container protocol -----------------------------------------------------------------------------------
    def __len__(self):
        return self.vector_size

    def _resolve(self, key):
        try:
            idx = int(key)
        except (TypeError, ValueError):
            return key
        return self.get_parameter_names()[idx]

    def __getitem__(self, key):
        return self.get_parameter(self._resolve(key))

    def __setitem__(self, key, value):
        return self.set_parameter(self._resolve(key), value)

    # -- sizes ------------------------------------------------------------------------------------------------
    @property
    def full_size(self):
        return len(self.parameter_names)

    @property
    def vector_size(self):
        return int(np.sum(self.unfrozen_mask))

    # -- the full vector (frozen parameters included) -----------------------------------------------------------
    @property
    def parameter_vector(self):
        return np.array([getattr(self, k) for k in self.parameter_names])

    @parameter_vector.setter
    def parameter_vector(self, v):
        if len(v) != self.full_size:
            raise ValueError("dimension mismatch")
        for k, val in zip(self.parameter_names, v):
            setattr(self, k, float(val))
        self.dirty = True

    def _select(self, seq, include_frozen):
        if include_frozen:
            return seq
        return [s for s, keep in zip(seq, self.unfrozen_mask) if keep]

    def get_parameter_names(self, include_frozen=False):
        if include_frozen:
            return self.parameter_names
        return tuple(self._select(self.parameter_names, False))

    def get_parameter_bounds(self, include_frozen=False):
        if include_frozen:
            return self.parameter_bounds
        return list(self._select(self.parameter_bounds, False))

    def get_parameter_vector(self, include_frozen=False):
        v = self.parameter_vector
        return v if include_frozen else v[self.unfrozen_mask]

    def get_parameter_dict(self, include_frozen=False):
        return OrderedDict(zip(self.get_parameter_names(include_frozen=include_frozen),
                               self.get_parameter_vector(include_frozen=include_frozen)))

    def set_parameter_vector(self, vector, include_frozen=False):
        v = self.parameter_vector
        if include_frozen:
            v[:] = vector
        else:
            v[self.unfrozen_mask] = vector
        self.parameter_vector = v
        self.dirty = True

    def check_parameter_vector(self, vector):
        """Would ``vector`` have a finite prior?  Leaves the model (and its dirty flag) untouched."""
        saved, was_dirty = np.array(self.get_parameter_vector()), self.dirty
        self.set_parameter_vector(vector)
        lp = self.log_prior()
        self.set_parameter_vector(saved)
        self.dirty = was_dirty
        return np.isfinite(lp)

    # -- freezing / thawing / single-parameter access ---------------------------------------------------------
    def _index_of(self, name):
        return self.get_parameter_names(include_frozen=True).index(name)

    def freeze_parameter(self, name):
        self.unfrozen_mask[self._index_of(name)] = False

    def thaw_parameter(self, name):
        self.unfrozen_mask[self._index_of(name)] = True

    def freeze_all_parameters(self):
        self.unfrozen_mask[:] = False

    def thaw_all_parameters(self):
        self.unfrozen_mask[:] = True

    def get_parameter(self, name):
        return self.get_parameter_vector(include_frozen=True)[self._index_of(name)]

    def set_parameter(self, name, value):
        v = self.get_parameter_vector(include_frozen=True)
        v[self._index_of(name)] = value
        self.set_parameter_vector(v, include_frozen=True)

    def log_prior(self):
        """0 inside the bounds box, -inf outside."""
        for p, (lo, hi) in zip(self.parameter_vector, self.parameter_bounds):
            if (lo is not None and p < lo) or (hi is not None and p > hi):
                return -np.inf
        return 0.0

    @staticmethod
    def parameter_sort(f):
        """Decorator: turn a ``{name: value}`` result into a list ordered like ``parameter_names``."""
        def wrapped(self, *args, **kwargs):
            by_name = f(self, *args, **kwargs)
            ordered = [by_name[k] for k in self.get_parameter_names(include_frozen=True)]
            if ordered and type(ordered[0]).__module__ == np.__name__:
                return np.vstack(ordered)
            return ordered
        return wrapped


class ModelSet(Model):
    """An ordered collection of named sub-models whose parameters are addressed as ``"name:parameter"``
    (a sub-model registered under ``None`` contributes its names unprefixed)."""

    def __init__(self, models):
        self.models = OrderedDict((name, model) for name, model in models)

    def __getattr__(self, name):
        if "models" in self.__dict__ and name in self.models:
            return self.models[name]
        raise AttributeError(name)

    @property
    def dirty(self):
        return any(m.dirty for m in self.models.values())

    @dirty.setter
    def dirty(self, value):
        for m in self.models.values():
            m.dirty = value

    @property
    def full_size(self):
        return sum(m.full_size for m in self.models.values())

    @property
    def vector_size(self):
        return sum(m.vector_size for m in self.models.values())

    @property
    def unfrozen_mask(self):
        return np.concatenate([m.unfrozen_mask for m in self.models.values()])

    @property
    def parameter_vector(self):
        return np.concatenate([m.parameter_vector for m in self.models.values()])

    @parameter_vector.setter
    def parameter_vector(self, v):
        pos = 0
        for m in self.models.values():
            n = m.full_size
            m.parameter_vector = v[pos:pos + n]
            pos += n

    @property
    def parameter_names(self):
        names = []
        for prefix, m in self.models.items():
            for k in m.parameter_names:
                names.append("{0}".format(k) if prefix is None else "{0}:{1}".format(prefix, k))
        return tuple(names)

    @property
    def parameter_bounds(self):
        out = []
        for m in self.models.values():
            out.extend(m.parameter_bounds)
        return out

    def _dispatch(self, method, name, *args):
        parts = name.split(":")
        key = parts[0]
        if key not in self.models:
            if None not in self.models:
                raise ValueError("unrecognized parameter '{0}'".format(name))
            key, parts = None, [None] + parts
        return getattr(self.models[key], method)(":".join(parts[1:]), *args)

    def freeze_parameter(self, name):
        self._dispatch("freeze_parameter", name)

    def thaw_parameter(self, name):
        self._dispatch("thaw_parameter", name)

    def freeze_all_parameters(self):
        for m in self.models.values():
            m.freeze_all_parameters()

    def thaw_all_parameters(self):
        for m in self.models.values():
            m.thaw_all_parameters()

    def get_parameter(self, name):
        return self._dispatch("get_parameter", name)

    def set_parameter(self, name, value):
        self.dirty = True
        return self._dispatch("set_parameter", name, value)

    def log_prior(self):
        total = 0.0
        for m in self.models.values():
            total += m.log_prior()
            if not np.isfinite(total):
                return -np.inf
        return total


class ConstantModel(Model):
    """``f(x) = value``."""

    parameter_names = ("value", )

    def get_value(self, x):
        return self.value + np.zeros(len(x))

    def compute_gradient(self, x):
        return np.ones((1, len(x)))


class CallableModel(Model):
    """Wrap a plain function (and optionally its gradient) as a parameter-free model."""

    def __init__(self, function, gradient=None):
        self.function = function
        self.gradient = gradient
        super(CallableModel, self).__init__()

    def get_value(self, x):
        return self.function(x)

    def compute_gradient(self, x):
        if self.gradient is not None:
            return self.gradient(x)
        return super(CallableModel, self).compute_gradient(x)

"""Functor handles: what the reference passes as Python callables becomes (name, packed parameters) here.

The reference's filters take ``jax``-traceable callables (``mfs/one_dim/filtering.py:41-47``).  A CUDA kernel cannot
call back into Python, so the callables are *named device functors* compiled into ``libmfs_b200.so``; these handles
only carry the ids and parameters across the C ABI.  Passing anything else to a filter raises ``TypeError`` -- there is
no CPU fallback.

Parameters may be scalars (shared by all filters) or arrays (one value per filter, e.g. a theta grid): they broadcast
against the batch axes of ``ys``.
"""
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

from . import _lib


@dataclass(frozen=True, eq=False)
class Drift:
    """Drift a(x) of dX = a(X) dt + b dW.  ``name`` in {'benes', 'well', 'linear'}."""
    name: str
    params: Tuple = ()

    def __post_init__(self):
        if self.name not in _lib.DRIFT:
            raise ValueError(f'unknown drift {self.name!r}; registered: {sorted(_lib.DRIFT)}')
        if len(self.params) > _lib.MAX_PARAMS:
            raise ValueError('too many drift parameters')


@dataclass(frozen=True, eq=False)
class Dispersion:
    """Constant dispersion b (every 1D model of the reference: ``mfs/one_dim/ss_models.py:39,76``)."""
    value: float = 1.


def benes_drift() -> Drift:
    """tanh(x)  (``mfs/one_dim/ss_models.py:37``)."""
    return Drift('benes')


def well_drift(theta1) -> Drift:
    """x (1 - theta1 x^2)  (``mfs/one_dim/ss_models.py:71``); ``theta1`` scalar or per-filter array."""
    return Drift('well', (theta1,))


def linear_drift(a) -> Drift:
    """a x  (OU: a = -1/ell, ``tests/test_filtering.py:45``)."""
    return Drift('linear', (a,))


@dataclass(frozen=True, eq=False)
class DriftND:
    """d-dimensional drift.  Registered: 'lotka_volterra' x * (x[::-1] * [-beta, delta] + [alpha, -gamma])
    (``mfs/multi_dims/ss_models.py:55-56``), params (alpha, beta, delta, gamma)."""
    name: str
    params: Tuple = ()

    def __post_init__(self):
        if self.name not in ('lotka_volterra',):
            raise ValueError(f'unknown d-dimensional drift {self.name!r}; registered: lotka_volterra')


@dataclass(frozen=True, eq=False)
class DispersionND:
    """d-dimensional dispersion.  Registered: 'proportional' diag(sigma * x) (``ss_models.py:58-59``), params (sigma,)."""
    name: str
    params: Tuple = ()

    def __post_init__(self):
        if self.name not in ('proportional',):
            raise ValueError(f'unknown d-dimensional dispersion {self.name!r}; registered: proportional')


def lotka_volterra_drift(alpha, beta, delta, gamma) -> DriftND:
    return DriftND('lotka_volterra', (alpha, beta, delta, gamma))


def proportional_dispersion(sigma) -> DispersionND:
    return DispersionND('proportional', (sigma,))


@dataclass(frozen=True, eq=False)
class TransitionSpec:
    family: str          # 'tme' | 'tme_normal' | 'euler' | 'normal_affine'
    drift: Drift
    dispersion: float
    dt: float
    order: int = 3
    params: Tuple = ()   # family parameters that are not drift parameters (normal_affine: F, Sigma)

    def packed_params(self):
        return self.params if self.family == 'normal_affine' else self.drift.params


@dataclass(frozen=True, eq=False)
class TransitionFunctor:
    """One of the five callables a ``sde_cond_moments_*`` factory returns
    (raw, central, scaled, mean, mean_var -- ``mfs/one_dim/moments.py:178-179``)."""
    spec: TransitionSpec
    role: str            # 'raw' | 'central' | 'scaled' | 'mean' | 'mean_var'

    def __call__(self, *_, **__):
        raise TypeError('TransitionFunctor is a device-functor handle, not a host callable (no CPU path).')


@dataclass(frozen=True, eq=False)
class MeasurementFunctor:
    """p(y | x): 'bernoulli_logistic_cubic' {c0, c1}, 'poisson_softplus' {theta2}, 'gaussian' {h, r}."""
    name: str
    params: Tuple = field(default_factory=tuple)

    def __post_init__(self):
        if self.name not in _lib.MEAS:
            raise ValueError(f'unknown measurement model {self.name!r}; registered: {sorted(_lib.MEAS)}')

    def __call__(self, *_, **__):
        raise TypeError('MeasurementFunctor is a device-functor handle, not a host callable (no CPU path).')

    def with_params(self, *params) -> 'MeasurementFunctor':
        return MeasurementFunctor(self.name, tuple(params))


def bernoulli_logistic_cubic(c0=5., c1=0.) -> MeasurementFunctor:
    """Bernoulli(1 / (1 + exp(-(x^3 / c0 - c1))))  (Benes: ``mfs/one_dim/ss_models.py:43-47``)."""
    return MeasurementFunctor('bernoulli_logistic_cubic', (c0, c1))


def poisson_softplus(theta2) -> MeasurementFunctor:
    """Poisson(log(1 + exp(theta2 x)))  (``mfs/one_dim/ss_models.py:80-84``)."""
    return MeasurementFunctor('poisson_softplus', (theta2,))


def gaussian(h=1., r=1.) -> MeasurementFunctor:
    """N(y; h x, r^2)  (``tests/test_filtering.py:41-42``)."""
    return MeasurementFunctor('gaussian', (h, r))


def pack_params(params, batch_shape, width: int = None) -> Tuple[np.ndarray, int]:
    """Pack scalars/arrays into a (B or 1, width) float64 table (width = MAX_PARAMS by default); returns
    (table, stride)."""
    width = width or _lib.MAX_PARAMS
    B = int(np.prod(batch_shape)) if len(batch_shape) else 1
    per_filter = any(np.ndim(p) > 0 for p in params)
    rows = B if per_filter else 1
    table = np.zeros((rows, width), dtype=np.float64)
    for k, p in enumerate(params):
        if np.ndim(p) > 0:
            table[:, k] = np.broadcast_to(np.asarray(p, dtype=np.float64), batch_shape).reshape(-1)
        else:
            table[:, k] = float(p)
    return table, (width if per_filter else 0)

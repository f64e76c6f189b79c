"""Host-side mirror of ``mfs/multi_dims/moments.py`` for the filter path: Gaussian product moments (Kan--Magnus
values, computed by the moment recursion) and the transition-moment factories returning device-functor handles."""
from dataclasses import dataclass
from typing import Tuple

import numpy as np

from ..functors import DriftND, DispersionND

__all__ = ['raw_moments_mvn_kan', 'central_moments_mvn_kan', 'sde_cond_moments_euler_maruyama',
           'sde_cond_moments_tme_normal', 'sde_cond_moments_tme', 'TransitionSpecND', 'TransitionFunctorND']


def _gaussian_product_moments(mean, cov, multi_indices):
    """E[prod_k X_k^{n_k}] for X ~ N(mean, cov) and every row of ``multi_indices`` by
    E[x^{n+e_i}] = mean_i E[x^n] + sum_j cov_ij n_j E[x^{n-e_j}]; equals Kan (2008), Prop. 2 (``moments.py:111-154``)."""
    mean, cov = np.asarray(mean, dtype=np.float64), np.asarray(cov, dtype=np.float64)
    d = mean.shape[0]
    table = {(0,) * d: 1.}

    def get(n):
        if n in table:
            return table[n]
        i = next(k for k in range(d) if n[k] > 0)
        base = list(n)
        base[i] -= 1
        val = mean[i] * get(tuple(base))
        for j in range(d):
            if base[j] > 0:
                low = list(base)
                low[j] -= 1
                val += cov[i, j] * base[j] * get(tuple(low))
        table[n] = val
        return val

    return np.array([get(tuple(int(v) for v in row)) for row in np.atleast_2d(multi_indices)])


def raw_moments_mvn_kan(mean, cov, multi_index) -> float:
    """E[X^n], X ~ N(mean, cov)  (``moments.py:111-154``)."""
    return float(_gaussian_product_moments(mean, cov, [multi_index])[0])


def central_moments_mvn_kan(cov, multi_index) -> float:
    """E[X^n], X ~ N(0, cov)  (``moments.py:66-108``)."""
    cov = np.asarray(cov, dtype=np.float64)
    return float(_gaussian_product_moments(np.zeros(cov.shape[0]), cov, [multi_index])[0])


@dataclass(frozen=True, eq=False)
class TransitionSpecND:
    family: str            # 'euler' | 'tme_normal' | 'tme'
    drift: DriftND
    dispersion: DispersionND
    dt: float
    order: int
    multi_indices: np.ndarray

    def packed_params(self):
        return tuple(self.drift.params) + tuple(self.dispersion.params)


@dataclass(frozen=True, eq=False)
class TransitionFunctorND:
    spec: TransitionSpecND
    role: str              # 'raw' | 'central' | 'scaled' | 'mean' | 'mean_var'

    def __call__(self, *_, **__):
        raise TypeError('TransitionFunctorND is a device-functor handle, not a host callable (no CPU path).')


_ROLES = ('raw', 'central', 'scaled', 'mean', 'mean_var')


def _factory(family, drift, dispersion, dt, order, multi_indices):
    if not isinstance(drift, DriftND) or not isinstance(dispersion, DispersionND):
        raise TypeError('drift / dispersion must be registered DriftND / DispersionND handles '
                        '(lotka_volterra_drift, proportional_dispersion): Python callables cannot run in the kernel.')
    spec = TransitionSpecND(family, drift, dispersion, float(dt), int(order), np.asarray(multi_indices))
    return tuple(TransitionFunctorND(spec, r) for r in _ROLES)


def sde_cond_moments_euler_maruyama(drift, dispersion, dt, multi_indices):
    """Mirror of ``mfs/multi_dims/moments.py:257-337``; use with the ``'index'`` signature flag."""
    return _factory('euler', drift, dispersion, dt, 1, multi_indices)


def sde_cond_moments_tme_normal(drift, dispersion, dt, tme_order, multi_indices):
    """Mirror of ``mfs/multi_dims/moments.py:340-411`` (orders 1 and 2); use with the ``'index'`` signature flag."""
    if tme_order not in (1, 2):
        raise ValueError('tme_order must be 1 or 2 for the d-dimensional TME-normal functor')
    return _factory('tme_normal', drift, dispersion, dt, tme_order, multi_indices)


def sde_cond_moments_tme(drift, dispersion, dt, tme_order):
    """Mirror of ``mfs/multi_dims/moments.py:414-479`` (TME without the Normal approximation, orders 1 and 2); use with
    the ``'multi-index'`` signature flag, like the reference (``dardel/prey_predator/mf.py``)."""
    if tme_order not in (1, 2):
        raise ValueError('tme_order must be 1 or 2 for the d-dimensional TME functor')
    return _factory('tme', drift, dispersion, dt, tme_order, None)

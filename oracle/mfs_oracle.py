"""
CPU oracle for the mfs moment-filter hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the reference algorithm (zgbkdlm/mfs, pure JAX).  It is
the checker that the CUDA path is compared against.  Nothing under ``mfs_b200/`` may import it; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do.

Why a restatement: the reference imports ``jax`` and ``tme`` at module scope, neither of which is
installed (nor installable, no network) in this image, so it cannot be executed directly.

Parity pinning (see DESIGN.md "Oracle"):
  * the filter recursions and the quadrature below are checked against golden vectors produced by
    executing the reference's *own source files* (``mfs/one_dim/filtering.py``,
    ``mfs/one_dim/quadtures.py``, ``mfs/one_dim/moments.py``) on a NumPy-backed ``jax`` shim
    (``tests/golden/make_golden.py``), and against the analytic known-answer tests of the reference's
    test-suite (Kalman filter, closed-form integrals, exact OU moments);
  * the TME transition moments restate the published definition of the un-vendored third-party
    package ``tme`` (``tme>=0.1.5``, ``requirements.txt:6``):  E[phi(X_dt)|x] ~= sum_r dt^r/r! (A^r phi)(x),
    A phi = a phi' + 1/2 b^2 phi''.  The reference holds no stored TME values for a nonlinear drift, so
    parity with ``tme`` itself for tanh / cubic drifts is UNPINNED beyond the reference's exact-OU tests
    (``tests/test_one_dim_moments.py:90-118``), which are replayed in ``tests/test_oracle_moments.py``.

All ``file:line`` citations are relative to the reference root.
"""
from __future__ import annotations

import math
import warnings
from functools import lru_cache
from typing import Callable, Sequence, Tuple

import numpy as np
import scipy.linalg
import scipy.special

__all__ = [
    'hankel_indices', 'ldl', 'ldl_chol', 'moment_quadrature',
    'moment_filter_rms', 'moment_filter_cms', 'moment_filter_scms',
    'raw_moment_of_normal', 'raw_moment_of_standard_normal', 'central_moment_of_normal',
    'raw_to_central', 'central_to_raw', 'raw_to_scaled', 'scaled_to_central',
    'gaussian_sum_1d', 'tme_1d', 'sde_cond_moments_tme', 'sde_cond_moments_tme_normal',
    'sde_cond_moments_euler', 'sde_cond_moments_normal_exact_ou',
    'benes_bernoulli', 'well_poisson', 'bernoulli_pmf', 'poisson_pmf', 'norm_pdf',
    'kalman_filter_1d', 'characteristic_fn',
]


# ----------------------------------------------------------------------------------------------------------------------
# Linear algebra: mfs/one_dim/quadtures.py:29-60, 83-133 ; mfs/utils.py:495-538
# ----------------------------------------------------------------------------------------------------------------------
def hankel_indices(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Index tables of the moment Hankel pair G[i,j]=ms[i+j], H[i,j]=ms[i+j+1]  (quadtures.py:29-60)."""
    inds = np.arange(n)[:, None] + np.arange(n)[None, :]
    return inds, inds + 1


def ldl(mat: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """LDL^T without pivoting, same loop as mfs/utils.py:495-523."""
    n = mat.shape[0]
    l = np.eye(n)
    l[1:, 0] = mat[1:, 0] / mat[0, 0]
    d = np.ones((n,)) * mat[0, 0]
    for j in range(1, n):
        v = l[j, :j] * d[:j]
        _d = mat[j, j] - np.dot(l[j, :j], v)
        d[j] = _d
        l[j + 1:, j] = (mat[j + 1:, j] - l[j + 1:, :j] @ v) / _d
    return l, d


def ldl_chol(mat: np.ndarray, eps: float = None) -> np.ndarray:
    """Modified Cholesky factor L diag(where(d<0, eps, sqrt(d)))  (mfs/utils.py:526-538)."""
    if eps is None:
        eps = 1e-8 * np.linalg.norm(mat, 'fro')
    l, d = ldl(mat)
    with np.errstate(invalid='ignore'):
        return np.einsum('ij,j->ij', l, np.where(d < 0, eps, np.sqrt(d)))


def _cholesky_nan(G: np.ndarray) -> np.ndarray:
    """jax.lax.linalg.cholesky semantics: lower factor, all-NaN (no exception) when G is not PD or not finite."""
    if not np.all(np.isfinite(G)):
        return np.full_like(G, np.nan)
    try:
        return scipy.linalg.cholesky(G, lower=True, check_finite=False)
    except scipy.linalg.LinAlgError:
        return np.full_like(G, np.nan)


def moment_quadrature(ms: np.ndarray, mean: float = 0., scale: float = 1., sort_nodes: bool = False,
                      ldl: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Gauss quadrature (weights, nodes) from 2n moments  (mfs/one_dim/quadtures.py:83-133).

    G = ms[i+j], H = ms[i+j+1]; R = chol(G) (lower); K = R^-1 H R^-T (two triangular solves); dense symmetric
    ``eigh`` of K (LAPACK ``syevd``, which is what XLA:CPU dispatches to; ``jax.lax.linalg.eigh`` symmetrises its
    input first); weights = V[0,:]**2, nodes = scale*lambda + mean.
    """
    ms = np.asarray(ms, dtype=np.float64)
    n = math.floor(ms.shape[0] / 2)                                         # quadtures.py:122
    G_inds, H_inds = hankel_indices(n)
    G, H = ms[G_inds], ms[H_inds]                                           # :125
    R = ldl_chol(G) if ldl else _cholesky_nan(G)                            # :127
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(H))):
        return np.full((n,), np.nan), np.full((n,), np.nan)
    with np.errstate(all='ignore'):
        try:
            Y = scipy.linalg.solve_triangular(R, H, lower=True, check_finite=False)        # R^-1 H
            K = scipy.linalg.solve_triangular(R, Y.T, lower=True, check_finite=False).T   # (R^-1 H) R^-T   :128-129
        except scipy.linalg.LinAlgError:     # exactly singular factor: XLA's trsm divides by zero -> inf/NaN
            return np.full((n,), np.nan), np.full((n,), np.nan)
    if not np.all(np.isfinite(K)):
        return np.full((n,), np.nan), np.full((n,), np.nan)
    K = 0.5 * (K + K.T)                                                     # eigh(symmetrize_input=True)
    eigen_vals, eigen_vectors = scipy.linalg.eigh(K, driver='evd', check_finite=False)  # :131
    weights, nodes = eigen_vectors[0, :] ** 2, scale * eigen_vals + mean    # :133
    if sort_nodes:
        order = np.argsort(nodes)
        weights, nodes = weights[order], nodes[order]
    return weights, nodes


# ----------------------------------------------------------------------------------------------------------------------
# Filters: mfs/one_dim/filtering.py
# ----------------------------------------------------------------------------------------------------------------------
def _pow(x, n):
    """jnp.power(float64, int): libm pow, 0**0 = 1."""
    with np.errstate(all='ignore'):
        return np.power(np.asarray(x, dtype=np.float64)[..., None], np.asarray(n)[None, :])


def moment_filter_rms(state_cond_raw_moments: Callable, measurement_cond_pdf: Callable, rms0, ys,
                      stable: bool = False):
    """Raw-moment filter  (mfs/one_dim/filtering.py:32-89).  Returns (rmss (T, 2N), nell)."""
    rms = np.asarray(rms0, dtype=np.float64).copy()
    num_moments = rms.shape[0]
    moment_powers = np.arange(num_moments)
    if num_moments % 2 != 0:
        warnings.warn(f'The order of moments {num_moments - 1} is not odd.')        # :65-66
    T = len(ys)
    rmss = np.empty((T, num_moments))
    nell = 0.
    with np.errstate(all='ignore'):
        for t in range(T):
            y = ys[t]
            weights, nodes = moment_quadrature(rms, sort_nodes=False, ldl=stable)              # :78
            rms = np.einsum('ij,i->j', state_cond_raw_moments(nodes, moment_powers), weights)  # :79
            weights, nodes = moment_quadrature(rms, sort_nodes=False, ldl=stable)              # :82
            lik = measurement_cond_pdf(y, nodes)
            pdf_y = np.dot(lik, weights)                                                       # :83
            rms = np.einsum('ij,i->j', _pow(nodes, moment_powers) * lik[:, None], weights) / pdf_y   # :84
            nell -= np.log(pdf_y)                                                              # :85
            rmss[t] = rms
    return rmss, nell


def moment_filter_cms(state_cond_central_moments: Callable, state_cond_mean: Callable,
                      measurement_cond_pdf: Callable, cms0, mean0, ys, stable: bool = False):
    """Central-moment filter  (mfs/one_dim/filtering.py:92-161).  Returns (cmss (T,2N), means (T,), nell)."""
    cms = np.asarray(cms0, dtype=np.float64).copy()
    mean = float(mean0)
    num_moments = cms.shape[0]
    orders = np.arange(num_moments)
    if num_moments % 2 != 0:
        warnings.warn(f'The order of moments {num_moments - 1} is not odd.')
    T = len(ys)
    cmss, means = np.empty((T, num_moments)), np.empty((T,))
    nell = 0.
    with np.errstate(all='ignore'):
        for t in range(T):
            y = ys[t]
            weights, nodes = moment_quadrature(cms, mean, sort_nodes=False, ldl=stable)        # :145
            cond_means = state_cond_mean(nodes)
            mean = np.dot(cond_means, weights)                                                 # :147
            cms = np.einsum('ij,i->j', state_cond_central_moments(nodes, orders, mean), weights)   # :148
            weights, nodes = moment_quadrature(cms, mean, sort_nodes=False, ldl=stable)        # :151
            lik = measurement_cond_pdf(y, nodes)
            pdf_y = np.dot(lik, weights)                                                       # :152
            mean = np.dot(nodes * lik, weights) / pdf_y                                        # :153
            cms = np.einsum('ij,i->j', _pow(nodes - mean, orders) * lik[:, None], weights) / pdf_y  # :154-156
            nell -= np.log(pdf_y)
            cmss[t], means[t] = cms, mean
    return cmss, means, nell


def moment_filter_scms(state_cond_scaled_central_moments: Callable, state_cond_mean_var: Callable,
                       measurement_cond_pdf: Callable, scms0, mean0, scale0, ys, stable: bool = False):
    """Scaled-central-moment filter  (mfs/one_dim/filtering.py:164-240)."""
    scms = np.asarray(scms0, dtype=np.float64).copy()
    mean, scale = float(mean0), float(scale0)
    num_moments = scms.shape[0]
    orders = np.arange(num_moments)
    if num_moments % 2 != 0:
        warnings.warn(f'The order of moments {num_moments - 1} is not odd.')
    T = len(ys)
    scmss, means, scales = np.empty((T, num_moments)), np.empty((T,)), np.empty((T,))
    nell = 0.
    with np.errstate(all='ignore'):
        for t in range(T):
            y = ys[t]
            weights, nodes = moment_quadrature(scms, mean, scale, sort_nodes=False, ldl=stable)   # :222
            cond_means, cond_vars = state_cond_mean_var(nodes)
            mean, scale = np.dot(cond_means, weights), np.sqrt(np.dot(cond_vars, weights))        # :224
            scms = np.einsum('ij,i->j', state_cond_scaled_central_moments(nodes, orders, mean, scale), weights)
            weights, nodes = moment_quadrature(scms, mean, scale, sort_nodes=False, ldl=stable)   # :228
            lik = measurement_cond_pdf(y, nodes)
            pdf_y = np.dot(lik, weights)
            mean = np.dot(nodes * lik, weights) / pdf_y                                           # :230
            scale = np.sqrt(np.dot((nodes - mean) ** 2 * lik, weights) / pdf_y)                   # :231-232
            scms = np.einsum('ij,i->j', _pow((nodes - mean) / scale, orders) * lik[:, None], weights) / pdf_y
            nell -= np.log(pdf_y)
            scmss[t], means[t], scales[t] = scms, mean, scale
    return scmss, means, scales, nell


# ----------------------------------------------------------------------------------------------------------------------
# Moments of Normals and conversions: mfs/one_dim/moments.py:31-138
# ----------------------------------------------------------------------------------------------------------------------
def central_moment_of_normal(variance, p: int):
    """moments.py:31-38."""
    if p % 2 == 0:
        return np.sqrt(variance) ** p * scipy.special.factorial2(p - 1, exact=True) if p > 0 else 1.
    return 0.


def raw_moment_of_standard_normal(p: int) -> float:
    """moments.py:41-67."""
    if p % 2 == 0:
        return math.factorial(p) / (2 ** (p / 2) * math.factorial(int(p / 2)))
    return 0.


def raw_moment_of_normal(mean, variance, p: int):
    """Binomial sum of moments.py:70-74 (same term order; note NaN*0 = NaN for variance<0, as in the reference)."""
    mean = np.asarray(mean, dtype=np.float64)
    variance = np.asarray(variance, dtype=np.float64)
    with np.errstate(all='ignore'):
        return sum([math.comb(p, m) * mean ** m * variance ** ((p - m) / 2) * raw_moment_of_standard_normal(p - m)
                    for m in range(p + 1)])


def raw_to_central(rms: np.ndarray) -> np.ndarray:
    """moments.py:86-103."""
    rms = np.asarray(rms, dtype=np.float64)
    s = rms.shape[0]
    bn = scipy.linalg.pascal(s, kind='lower', exact=True).astype(np.float64)
    out = np.zeros(s)
    for n in range(s):
        for j in range(n + 1):
            out[n] += bn[n, j] * (-1.) ** (n - j) * rms[j] * rms[1] ** (n - j)
    return out


def central_to_raw(cms: np.ndarray, mean: float) -> np.ndarray:
    """moments.py:106-125."""
    cms = np.asarray(cms, dtype=np.float64)
    s = cms.shape[0]
    bn = scipy.linalg.pascal(s, kind='lower', exact=True).astype(np.float64)
    out = np.zeros(s)
    for n in range(s):
        for j in range(n + 1):
            out[n] += bn[n, j] * cms[j] * mean ** (n - j)
    return out


def raw_to_scaled(rms: np.ndarray, scale: float = None) -> np.ndarray:
    """moments.py:128-134."""
    rms = np.asarray(rms, dtype=np.float64)
    if scale is None:
        scale = np.sqrt(rms[2] - rms[1] ** 2)
    return raw_to_central(rms) / np.array([scale ** n for n in range(rms.shape[0])])


def scaled_to_central(sms: np.ndarray, scale: float) -> np.ndarray:
    """moments.py:137-140."""
    return np.asarray(sms) * np.array([scale ** n for n in range(len(sms))])


class GaussianSum1D:
    """Moments of a 1D Gaussian mixture  (mfs/utils.py:39-74, ``GaussianSum1D.new``)."""

    def __init__(self, means, variances, weights, N: int = 2):
        self.means, self.variances, self.weights = map(lambda a: np.asarray(a, dtype=np.float64),
                                                       (means, variances, weights))
        centre = np.sum(self.means * self.weights)
        self.rms = np.array([sum([raw_moment_of_normal(m, v, p) * w for m, v, w in
                                  zip(self.means, self.variances, self.weights)]) for p in range(2 * N)])
        self.cms = np.array([sum([raw_moment_of_normal(m - centre, v, p) * w for m, v, w in
                                  zip(self.means, self.variances, self.weights)]) for p in range(2 * N)])
        self.mean = centre
        self.variance = self.cms[2]
        self.scms = self.cms / np.sqrt(self.variance) ** np.arange(2 * N)


def gaussian_sum_1d(means, variances, weights, N: int = 2) -> GaussianSum1D:
    return GaussianSum1D(means, variances, weights, N)


# ----------------------------------------------------------------------------------------------------------------------
# TME (third-party ``tme`` package, restated from its definition)
# ----------------------------------------------------------------------------------------------------------------------
_DRIFTS = {
    # name -> (sympy expression builder of the drift, parameter names)
    'benes': (lambda sp, u, p: sp.tanh(u), ()),                               # ss_models.py:37
    'well': (lambda sp, u, p: u * (1 - p[0] * u ** 2), ('theta1',)),          # ss_models.py:71
    'ou': (lambda sp, u, p: -u / p[0], ('ell',)),                             # tests/test_filtering.py:45-46
    'linear': (lambda sp, u, p: p[0] * u, ('a',)),                            # tests/test_one_dim_moments.py:61-62
}


@lru_cache(maxsize=None)
def _tme_lambdas(drift_name: str, order: int, num_moments: int):
    """Build numpy callables for E[((X_dt - m)/s)^p | x], p=0..num_moments-1, straight from the TME definition
    (generator applied ``order`` times to each phi_p separately, like ``tme.expectation`` does by autodiff;
    call sites mfs/one_dim/moments.py:151,159,167,171)."""
    import sympy as sp
    u, m, s, dt, b = sp.symbols('u m s dt b', real=True)
    builder, pnames = _DRIFTS[drift_name]
    psyms = sp.symbols(' '.join(pnames) + ' _dummy', real=True)[:len(pnames)] if pnames else ()
    a = builder(sp, u, psyms)

    def gen(phi):
        return a * sp.diff(phi, u) + sp.Rational(1, 2) * b ** 2 * sp.diff(phi, u, 2)

    def expectation(phi):
        out, cur = phi, phi
        for r in range(1, order + 1):
            cur = gen(cur)
            out = out + dt ** r / math.factorial(r) * cur
        return out

    exprs = [expectation(((u - m) / s) ** p) for p in range(num_moments)]
    args = (u, m, s, dt, b) + tuple(psyms)
    fn_moments = sp.lambdify(args, exprs, modules='numpy', cse=True)

    # tme.mean_and_cov: mean = expansion of phi=x; cov = sum_{r>=1} dt^r/r! [A^r(x^2) - sum_s C(r,s) A^s x A^{r-s} x]
    pows_i = [u]
    pows_ii = [u ** 2]
    for r in range(1, order + 1):
        pows_i.append(gen(pows_i[-1]))
        pows_ii.append(gen(pows_ii[-1]))
    mean_expr = sum(dt ** r / math.factorial(r) * pows_i[r] for r in range(order + 1))
    var_expr = 0
    for r in range(1, order + 1):
        phi_r = pows_ii[r] - sum(math.comb(r, q) * pows_i[q] * pows_i[r - q] for q in range(r + 1))
        var_expr = var_expr + dt ** r / math.factorial(r) * phi_r
    fn_mean_var = sp.lambdify(args, [mean_expr, sp.expand(var_expr)], modules='numpy', cse=True)
    return fn_moments, fn_mean_var


def tme_1d(drift_name: str, params: Sequence[float], b: float, dt: float, order: int, num_moments: int):
    """Return (moments(x, m, s) -> (n, num_moments), mean_var(x) -> (mean, var)) for a named 1D drift with constant
    dispersion ``b``."""
    fn_moments, fn_mean_var = _tme_lambdas(drift_name, order, num_moments)
    params = tuple(float(p) for p in params)

    def moments(x, m=0., s=1.):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        with np.errstate(all='ignore'):
            cols = fn_moments(x, m, s, dt, b, *params)
        return np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), x.shape) for c in cols], axis=-1)

    def mean_var(x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        with np.errstate(all='ignore'):
            mu, var = fn_mean_var(x, 0., 1., dt, b, *params)
        return np.broadcast_to(np.asarray(mu, dtype=np.float64), x.shape).copy(), \
            np.broadcast_to(np.asarray(var, dtype=np.float64), x.shape).copy()

    return moments, mean_var


def _check_orders(orders, num_moments):
    orders = np.asarray(orders)
    if orders.max(initial=0) >= num_moments:
        raise ValueError('moment order out of range')
    return orders


def sde_cond_moments_tme(drift_name: str, params, b: float, dt: float, tme_order: int, num_moments: int = 32):
    """Five callables of mfs/one_dim/moments.py:141-179, for a named drift."""
    moments, mean_var = tme_1d(drift_name, params, b, dt, tme_order, num_moments)

    def state_cond_raw_moments(x, n):
        return moments(x)[..., _check_orders(n, num_moments)]

    def state_cond_central_moments(x, n, mean):
        return moments(x, mean)[..., _check_orders(n, num_moments)]

    def state_cond_scaled_central_moments(x, n, mean, scale):
        return moments(x, mean, scale)[..., _check_orders(n, num_moments)]

    def state_cond_mean(x):
        return moments(x)[..., 1]

    def state_cond_mean_var(x):
        return mean_var(x)

    return state_cond_raw_moments, state_cond_central_moments, state_cond_scaled_central_moments, state_cond_mean, \
        state_cond_mean_var


def _normal_family(mean_var: Callable, cond_mean: Callable, N: int):
    """Shared body of moments.py:182-219 and :222-255: moments of N(mu(x) - m, v(x)) via ``raw_moment_of_normal``."""
    num_moments = 2 * N

    def state_cond_raw_moments(x, n):
        mu, var = mean_var(x)
        return np.stack([raw_moment_of_normal(mu, var, p) * np.ones_like(mu) for p in range(num_moments)],
                        axis=-1)[..., np.asarray(n)]

    def state_cond_central_moments(x, n, mean):
        mu, var = mean_var(x)
        return np.stack([raw_moment_of_normal(mu - mean, var, p) * np.ones_like(mu) for p in range(num_moments)],
                        axis=-1)[..., np.asarray(n)]

    def state_cond_scaled_central_moments(x, n, mean, scale):
        # The reference divides EVERY order by prod(scale**arange(2N)) (moments.py:205, :243) -- restated as is.
        mu, var = mean_var(x)
        s = np.prod(scale ** np.arange(num_moments))
        return np.stack([raw_moment_of_normal(mu - mean, var, p) * np.ones_like(mu) for p in range(num_moments)],
                        axis=-1)[..., np.asarray(n)] / s

    return state_cond_raw_moments, state_cond_central_moments, state_cond_scaled_central_moments, cond_mean, mean_var


def sde_cond_moments_tme_normal(drift_name: str, params, b: float, dt: float, tme_order: int, N: int):
    """mfs/one_dim/moments.py:182-219."""
    moments, mean_var = tme_1d(drift_name, params, b, dt, tme_order, 2)
    return _normal_family(mean_var, lambda x: moments(x)[..., 1], N)


def sde_cond_moments_euler(drift_name: str, params, b: float, dt: float, N: int):
    """mfs/one_dim/moments.py:222-255 (Euler--Maruyama + Normal)."""
    import sympy as sp
    u = sp.symbols('u', real=True)
    builder, pnames = _DRIFTS[drift_name]
    a = sp.lambdify(u, builder(sp, u, tuple(params)), modules='numpy')

    def mean_var(x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        return x + a(x) * dt, np.full_like(x, b ** 2 * dt)

    return _normal_family(mean_var, lambda x: mean_var(x)[0], N)


def sde_cond_moments_normal_exact_ou(F: float, Sigma: float, N: int):
    """Exact OU transition N(F x, Sigma) of dardel/convergence/convergence_mf.py:86-107."""

    def mean_var(x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        return F * x, np.full_like(x, Sigma)

    return _normal_family(mean_var, lambda x: mean_var(x)[0], N)


# ----------------------------------------------------------------------------------------------------------------------
# Measurement models
# ----------------------------------------------------------------------------------------------------------------------
def bernoulli_pmf(y, p):
    """jax.scipy.stats.bernoulli.pmf = exp(xlogy(y,p) + xlog1py(1-y,-p))  (used at ss_models.py:46)."""
    p = np.asarray(p, dtype=np.float64)
    with np.errstate(all='ignore'):
        return np.exp(scipy.special.xlogy(float(y), p) + scipy.special.xlog1py(1. - float(y), -p))


def poisson_pmf(k, mu):
    """jax.scipy.stats.poisson.pmf = exp(xlogy(k,mu) - gammaln(k+1) - mu)  (used at ss_models.py:83)."""
    mu = np.asarray(mu, dtype=np.float64)
    with np.errstate(all='ignore'):
        return np.exp(scipy.special.xlogy(float(k), mu) - scipy.special.gammaln(float(k) + 1.) - mu)


def norm_pdf(y, loc, scale):
    """jax.scipy.stats.norm.pdf (tests/test_filtering.py:41-42)."""
    z = (y - np.asarray(loc, dtype=np.float64)) / scale
    return np.exp(-0.5 * z * z) / (scale * math.sqrt(2 * math.pi))


def benes_bernoulli(N: int = 2):
    """Constants and callables of mfs/one_dim/ss_models.py:25-56 (no simulator)."""
    dt, T = 1e-2, 100
    init_cond = gaussian_sum_1d([-0.5, 0.5], [0.05, 0.05], [0.5, 0.5], N)

    def logistic(x):
        with np.errstate(all='ignore'):
            return 1 / (1 + np.exp(-np.asarray(x, dtype=np.float64) ** 3 / 5))      # ss_models.py:43

    def measurement_cond_pmf(y, x):
        return bernoulli_pmf(y, logistic(x))                                       # :46

    return dt, T, init_cond, logistic, measurement_cond_pmf


def well_poisson(N: int = 2):
    """Constants and callables of mfs/one_dim/ss_models.py:59-93."""
    dt, T = 1e-2, 1000
    init_cond = gaussian_sum_1d([-0.5, 0.5], [0.05, 0.05], [0.5, 0.5], N)

    def emission(x, p):
        with np.errstate(all='ignore'):
            return np.log(1. + np.exp(p * np.asarray(x, dtype=np.float64)))         # :80

    def measurement_cond_pmf(y, x, p):
        return poisson_pmf(y, emission(x, p))                                      # :83

    return dt, T, init_cond, emission, measurement_cond_pmf


# ----------------------------------------------------------------------------------------------------------------------
# Analytic cross-checks used by the reference's own tests
# ----------------------------------------------------------------------------------------------------------------------
def kalman_filter_1d(F, Sigma, H, R, mean0, var0, ys):
    """Scalar Kalman filter  (tests/test_filtering.py:61-77 with H=1)."""
    mf, vf, nell = mean0, var0, 0.
    mfs, vfs = np.empty(len(ys)), np.empty(len(ys))
    for t, y in enumerate(ys):
        mp, vp = F * mf, F * vf * F + Sigma
        s = H * vp * H + R
        k = vp * H / s
        mf, vf = mp + k * (y - H * mp), vp - k * H * vp
        nell -= -0.5 * math.log(2 * math.pi * s) - 0.5 * (y - H * mp) ** 2 / s
        mfs[t], vfs[t] = mf, vf
    return mfs, vfs, nell


def characteristic_fn(z, ms, mean=0., scale=1.):
    """mfs/one_dim/moments.py:309-337."""
    weights, nodes = moment_quadrature(ms, mean, scale, sort_nodes=False)
    return np.dot(np.exp(1.j * z * nodes), weights)

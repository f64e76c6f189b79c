"""
CPU oracle for the multi-dimensional moment filter  --  TEST INFRASTRUCTURE ONLY (see oracle/mfs_oracle.py header).

NumPy/SciPy restatement of
  moment_quadrature_nd      mfs/multi_dims/quadratures.py:120-178 (+ nd_cartesian_prod_indices :29-48)
  moment_filter_nd_rms/cms  mfs/multi_dims/filtering.py:283-344 / 210-280
  raw_moments_mvn_kan       mfs/multi_dims/moments.py:111-154 (Kan 2008, Prop. 2)
  nd transition factories   mfs/multi_dims/moments.py:257-411 (Euler--Maruyama, TME + Normal)
  prey_predator             mfs/multi_dims/ss_models.py:40-67 ; GaussianSumND.new mfs/utils.py:113-131
TME mean/covariance (third-party ``tme.mean_and_cov``, absent) is restated from its definition with sympy.
Pinned by ``tests/golden/golden_nd_*.npz`` (reference source executed on the NumPy jax-shim).
"""
import itertools
import math
from functools import lru_cache

import numpy as np
import scipy.linalg
from scipy.special import comb as _comb, factorial as _factorial

from . import mfs_oracle as O1


def nd_cartesian_prod_indices(d, n):
    """quadratures.py:29-48."""
    return np.asarray(tuple(itertools.product(*([list(range(n))] * d))), dtype='int64')


def moment_quadrature_nd(ms, inds, mean=None, scale=None, ldl=False):
    """quadratures.py:120-178: one Cholesky of the Gram matrix, d matrices K_i = R^-1 H_i R^-T, d dense symmetric
    eigen-decompositions, Cartesian product of the eigenpairs, weights = prod_k <v_k, v_{k+1}> * v_0[0] * v_{d-1}[0]."""
    ms = np.asarray(ms, dtype=np.float64)
    d, n = inds.shape[0] - 1, inds.shape[1]
    nan_out = (np.full((n ** d,), np.nan), np.full((n ** d, d), np.nan))
    G, Hs = ms[inds[0]], ms[inds[1:]]
    R = O1.ldl_chol(G) if ldl else O1._cholesky_nan(G)
    if not np.all(np.isfinite(R)) or not np.all(np.isfinite(Hs)):
        return nan_out
    eigvectors, eigvals = np.empty((d, n, n)), np.empty((d, n))
    with np.errstate(all='ignore'):
        for k in range(d):
            Y = scipy.linalg.solve_triangular(R, Hs[k], lower=True, check_finite=False)
            K = scipy.linalg.solve_triangular(R, Y.T, lower=True, check_finite=False).T
            if not np.all(np.isfinite(K)):
                return nan_out
            K = 0.5 * (K + K.T)
            eigvals[k], eigvectors[k] = scipy.linalg.eigh(K, driver='evd', check_finite=False)
    combs = nd_cartesian_prod_indices(d, n)
    nodes = np.stack([eigvals[k][combs[:, k]] for k in range(d)], axis=1)
    vecs = [eigvectors[k][:, combs[:, k]] for k in range(d)]                 # each (n, n**d)
    weights = np.ones(n ** d)
    for k in range(d - 1):
        weights = weights * np.einsum('ic,ic->c', vecs[k], vecs[k + 1])
    weights = weights * vecs[0][0] * vecs[-1][0]
    if mean is None:
        return weights, nodes
    if scale is None:
        return weights, nodes + np.asarray(mean)
    return weights, nodes * np.asarray(scale) + np.asarray(mean)


def _powers(x, multi_indices):
    """prod_k x_k ** n_k for all nodes (r, d) and multi-indices (z, d) -> (r, z)."""
    with np.errstate(all='ignore'):
        return np.prod(np.power(x[:, None, :], multi_indices[None, :, :]), axis=-1)


def moment_filter_nd_rms(state_cond_raw_moments, measurement_cond_pdf, ys, moments_partial_order, rms0, stable=False):
    """filtering.py:283-344.  ``state_cond_raw_moments(nodes) -> (r, z)``."""
    multi_indices, inds = moments_partial_order
    rms = np.asarray(rms0, dtype=np.float64).copy()
    T = len(ys)
    rmss = np.empty((T, rms.shape[0]))
    nell = 0.
    with np.errstate(all='ignore'):
        for t in range(T):
            w, x = moment_quadrature_nd(rms, inds, ldl=stable)
            rms = np.einsum('ij,i->j', state_cond_raw_moments(x), w)
            w, x = moment_quadrature_nd(rms, inds, ldl=stable)
            lik = measurement_cond_pdf(ys[t], x)
            pdf_y = np.dot(lik, w)
            rms = np.einsum('ij,i->j', _powers(x, multi_indices) * lik[:, None], w) / pdf_y
            nell -= np.log(pdf_y)
            rmss[t] = rms
    return rmss, nell


def moment_filter_nd_cms(state_cond_central_moments, state_cond_mean, measurement_cond_pdf, ys, moments_partial_order,
                         cms0, mean0, stable=False):
    """filtering.py:210-280.  ``state_cond_central_moments(nodes, mean) -> (r, z)``, ``state_cond_mean(nodes) -> (r, d)``."""
    multi_indices, inds = moments_partial_order
    cms = np.asarray(cms0, dtype=np.float64).copy()
    mean = np.asarray(mean0, dtype=np.float64).copy()
    T = len(ys)
    cmss, means = np.empty((T, cms.shape[0])), np.empty((T, mean.shape[0]))
    nell = 0.
    with np.errstate(all='ignore'):
        for t in range(T):
            w, x = moment_quadrature_nd(cms, inds, mean, ldl=stable)
            mean = np.einsum('ij,i->j', state_cond_mean(x), w)
            cms = np.einsum('ij,i->j', state_cond_central_moments(x, mean), w)
            w, x = moment_quadrature_nd(cms, inds, mean, ldl=stable)
            lik = measurement_cond_pdf(ys[t], x)
            pdf_y = np.dot(lik, w)
            mean = np.einsum('ij,i->j', x * lik[:, None], w) / pdf_y
            cms = np.einsum('ij,i->j', _powers(x - mean, multi_indices) * lik[:, None], w) / pdf_y
            nell -= np.log(pdf_y)
            cmss[t], means[t] = cms, mean
    return cmss, means, nell


def raw_moments_mvn_kan(mean, cov, multi_index):
    """Kan (2008) Proposition 2, same summation as moments.py:111-154."""
    multi_index = np.asarray(multi_index)
    s = int(multi_index.sum())
    ranges = [tuple(range(int(sn) + 1)) for sn in multi_index] + [tuple(range(s // 2 + 1))]
    vs_and_r = np.asarray(tuple(itertools.product(*ranges)), dtype='int64')
    vs, rs = vs_and_r[:, :-1], vs_and_r[:, -1]
    hs = multi_index / 2 - vs
    signs = (-1.) ** np.sum(vs, axis=1)
    combs = np.prod(_comb(multi_index, vs), axis=1)
    with np.errstate(all='ignore'):
        quad = (np.einsum('ci,ij,cj->c', hs, cov, hs) / 2) ** rs * (hs @ mean) ** (s - 2 * rs) \
            / (_factorial(rs, exact=False) * _factorial(s - 2 * rs, exact=False))
    return float(np.einsum('i,i,i', signs, combs, quad))


def gaussian_moments_recursive(mean, cov, multi_indices):
    """All product moments E[prod X_k^{n_k}], X ~ N(mean, cov), by the recursion
    E[x^{n+e_i}] = mean_i E[x^n] + sum_j cov_ij n_j E[x^{n-e_j}]  (equals Kan's sum; used to cross-check it)."""
    table = {}
    d = len(mean)
    for n in sorted((tuple(int(v) for v in r) for r in multi_indices), key=lambda r: (sum(r), r)):
        if sum(n) == 0:
            table[n] = 1.
            continue
        i = next(k for k in range(d) if n[k] > 0)
        base = list(n)
        base[i] -= 1
        val = mean[i] * _lookup(table, tuple(base), mean, cov)
        for j in range(d):
            if base[j] > 0:
                lower = list(base)
                lower[j] -= 1
                val += cov[i][j] * base[j] * _lookup(table, tuple(lower), mean, cov)
        table[n] = val
    return np.array([table[tuple(int(v) for v in r)] for r in multi_indices])


def _lookup(table, n, mean, cov):
    if n not in table:
        table[n] = gaussian_moments_recursive(mean, cov, [n])[0] if sum(n) else 1.
    return table[n]


class GaussianSumND:
    """mfs/utils.py:77-131 (``new``)."""

    def __init__(self, means, covs, weights, multi_indices):
        means, covs, weights = (np.asarray(a, dtype=np.float64) for a in (means, covs, weights))
        self.d = means.shape[1]
        self.means, self.covs, self.weights = means, covs, weights
        centre = np.sum(means * weights[:, None], axis=0)
        self.mean = centre
        self.cov = sum(w * (c + np.outer(m, m)) for m, c, w in zip(means, covs, weights)) - np.outer(centre, centre)
        self.rms = sum(w * np.array([raw_moments_mvn_kan(m, c, n) for n in multi_indices])
                       for m, c, w in zip(means, covs, weights))
        self.cms = sum(w * np.array([raw_moments_mvn_kan(m - centre, c, n) for n in multi_indices])
                       for m, c, w in zip(means, covs, weights))


# ---- Lotka--Volterra (prey--predator) model: mfs/multi_dims/ss_models.py:40-67 ---------------------------------------
LV = dict(alp=4., beta=4., delta=4., gamma=4., sigma=0.1, dt=1e-3, T=2000)


def lv_drift(x, p=LV):
    x = np.asarray(x, dtype=np.float64)
    return x * (x[..., ::-1] * np.array([-p['beta'], p['delta']]) + np.array([p['alp'], -p['gamma']]))


def lv_emission(x):
    with np.errstate(all='ignore'):
        return 1 / (1 + np.exp(-np.asarray(x, dtype=np.float64) ** 3 + 1))


def lv_measurement_pmf(y, x):
    """Bernoulli on x[0] (ss_models.py:63-67)."""
    return O1.bernoulli_pmf(y, lv_emission(np.asarray(x)[..., 0]))


def prey_predator(multi_indices):
    gs = GaussianSumND(np.array([[1., 1.], [1., 1.]]), np.array([np.eye(2), 2 * np.eye(2)]) * 0.001,
                       np.array([0.5, 0.5]), multi_indices)
    return LV['dt'], LV['T'], gs


@lru_cache(maxsize=None)
def _lv_tme_mean_cov(order):
    """tme.mean_and_cov for the Lotka--Volterra SDE from the TME definition (generator applied symbolically)."""
    import sympy as sp
    x1, x2, dt, al, be, de, ga, sg = sp.symbols('x1 x2 dt alp beta delta gamma sigma', real=True)
    X = [x1, x2]
    a = [x1 * (al - be * x2), x2 * (de * x1 - ga)]
    bbT = [[sg ** 2 * x1 ** 2, 0], [0, sg ** 2 * x2 ** 2]]

    def gen(phi):
        return sum(a[i] * sp.diff(phi, X[i]) for i in range(2)) + sp.Rational(1, 2) * sum(
            bbT[i][j] * sp.diff(phi, X[i], X[j]) for i in range(2) for j in range(2))

    pw1 = [[x1], [x2]]
    pw2 = {(i, j): [X[i] * X[j]] for i in range(2) for j in range(2)}
    for r in range(1, order + 1):
        for i in range(2):
            pw1[i].append(sp.expand(gen(pw1[i][-1])))
        for key in pw2:
            pw2[key].append(sp.expand(gen(pw2[key][-1])))
    mean = [sum(dt ** r / math.factorial(r) * pw1[i][r] for r in range(order + 1)) for i in range(2)]
    cov = {}
    for (i, j) in pw2:
        acc = 0
        for r in range(1, order + 1):
            phi_r = pw2[(i, j)][r] - sum(math.comb(r, s) * pw1[i][s] * pw1[j][r - s] for s in range(r + 1))
            acc = acc + dt ** r / math.factorial(r) * phi_r
        cov[(i, j)] = sp.expand(acc)
    args = (x1, x2, dt, al, be, de, ga, sg)
    return sp.lambdify(args, mean + [cov[(0, 0)], cov[(0, 1)], cov[(1, 1)]], modules='numpy', cse=True)


def lv_mean_cov(x, kind, order=2, p=LV):
    """Conditional mean (r, 2) and covariance (r, 2, 2) for nodes x (r, 2).  kind: 'euler' | 'tme_normal'."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    if kind == 'euler':                                                      # moments.py:291-293
        mean = x + lv_drift(x, p) * p['dt']
        cov = np.zeros(x.shape[:-1] + (2, 2))
        cov[..., 0, 0] = (p['sigma'] * x[..., 0]) ** 2 * p['dt']
        cov[..., 1, 1] = (p['sigma'] * x[..., 1]) ** 2 * p['dt']
        return mean, cov
    fn = _lv_tme_mean_cov(order)
    m1, m2, c11, c12, c22 = fn(x[..., 0], x[..., 1], p['dt'], p['alp'], p['beta'], p['delta'], p['gamma'], p['sigma'])
    ones = np.ones(x.shape[:-1])
    mean = np.stack([m1 * ones, m2 * ones], axis=-1)
    cov = np.empty(x.shape[:-1] + (2, 2))
    cov[..., 0, 0], cov[..., 0, 1], cov[..., 1, 0], cov[..., 1, 1] = c11 * ones, c12 * ones, c12 * ones, c22 * ones
    return mean, cov


def lv_cond_moments(kind, multi_indices, order=2, use_kan=True):
    """(state_cond_raw_moments(x), state_cond_central_moments(x, mean), state_cond_mean(x)) for the LV model with the
    Normal transition families (moments.py:257-411)."""
    multi_indices = np.asarray(multi_indices)
    fn = raw_moments_mvn_kan if use_kan else None

    def moments(x, shift):
        mean, cov = lv_mean_cov(x, kind, order)
        out = np.empty((mean.shape[0], multi_indices.shape[0]))
        for r in range(mean.shape[0]):
            if use_kan:
                out[r] = [fn(mean[r] - shift, cov[r], n) for n in multi_indices]
            else:
                out[r] = gaussian_moments_recursive(mean[r] - shift, cov[r], multi_indices)
        return out

    return (lambda x: moments(x, 0.)), (lambda x, m: moments(x, np.asarray(m))), (lambda x: lv_mean_cov(x, kind, order)[0])


@lru_cache(maxsize=None)
def _lv_tme_moment_fns(order, multi_indices_key):
    """tme.expectation(phi_n, x, dt, drift, dispersion, order) for phi_n(u) = prod((u - m)^n), every multi-index n
    (mfs/multi_dims/moments.py:441-459): the generator is applied symbolically to EACH monomial, `order` times."""
    import sympy as sp
    x1, x2, m1, m2, dt, al, be, de, ga, sg = sp.symbols('x1 x2 m1 m2 dt alp beta delta gamma sigma', real=True)
    X = [x1, x2]
    a = [x1 * (al - be * x2), x2 * (de * x1 - ga)]
    g = [sg ** 2 * x1 ** 2, sg ** 2 * x2 ** 2]

    def gen(phi):
        return sum(a[i] * sp.diff(phi, X[i]) + sp.Rational(1, 2) * g[i] * sp.diff(phi, X[i], 2) for i in range(2))

    exprs = []
    for n in multi_indices_key:
        cur = (x1 - m1) ** n[0] * (x2 - m2) ** n[1]
        tot = cur
        for r in range(1, order + 1):
            cur = gen(cur)
            tot = tot + dt ** r / math.factorial(r) * cur
        exprs.append(tot)
    return sp.lambdify((x1, x2, m1, m2, dt, al, be, de, ga, sg), exprs, modules='numpy', cse=True)


def lv_cond_moments_tme(multi_indices, order=2, p=LV):
    """(state_cond_raw_moments(x), state_cond_central_moments(x, mean), state_cond_mean(x)) of
    ``sde_cond_moments_tme`` (moments.py:414-479, 'multi-index' signature) for the LV model."""
    multi_indices = np.asarray(multi_indices)
    key = tuple(tuple(int(v) for v in row) for row in multi_indices)
    fn = _lv_tme_moment_fns(order, key)
    fn_mean = _lv_tme_moment_fns(order, ((1, 0), (0, 1)))

    def moments(x, shift):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        shift = np.broadcast_to(np.asarray(shift, dtype=np.float64), (2,))
        cols = fn(x[:, 0], x[:, 1], shift[0], shift[1], p['dt'], p['alp'], p['beta'], p['delta'], p['gamma'], p['sigma'])
        return np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), x.shape[:1]) for c in cols], axis=-1)

    def mean(x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        cols = fn_mean(x[:, 0], x[:, 1], 0., 0., p['dt'], p['alp'], p['beta'], p['delta'], p['gamma'], p['sigma'])
        return np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), x.shape[:1]) for c in cols], axis=-1)

    return (lambda x: moments(x, 0.)), (lambda x, m: moments(x, m)), mean

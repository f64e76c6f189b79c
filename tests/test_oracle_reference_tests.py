"""Replay of the reference's own analytic known-answer tests against the oracle (the reference stores no golden files;
its oracles are closed forms -- SURVEY.md section 4).  Each test cites the reference test it restates."""
import math

import numpy as np
import pytest

from oracle import mfs_oracle as O
from oracle import c_oracle as C


def _ou_test_data():
    """tests/test_filtering.py:18-37 (numpy seed 666 -> reproducible without JAX)."""
    np.random.seed(666)
    dt, T = 1e-2, 100
    ts = np.linspace(dt, dt * T, T)
    ell, sigma = 1., 0.5
    cov = np.exp(-np.abs(ts[None, :] - ts[:, None]) / ell) * sigma ** 2
    ys = np.linalg.cholesky(cov) @ np.random.randn(T) + np.random.randn(T)
    return dt, T, ell, sigma, ys


def test_1d_convergence_to_kalman():
    """tests/test_filtering.py:82-111: N=10, OU + Gaussian likelihood, TME-3; mean rtol 1e-2, var 1e-3, nell 1e-5."""
    dt, T, ell, sigma, ys = _ou_test_data()
    np.testing.assert_allclose(ys[:3], [1.02147222, 1.03211481, -0.11501503], atol=1e-8)   # SURVEY 8c probe
    b = math.sqrt(2) * sigma / math.sqrt(ell)
    N = 10
    f_rms = O.sde_cond_moments_tme('ou', (ell,), b, dt, 3, 2 * N)[0]
    rms0 = np.array([O.raw_moment_of_normal(0.1, 0.1, p) for p in range(2 * N)])
    rmss, nell = O.moment_filter_rms(f_rms, lambda y, x: O.norm_pdf(y, x, 1.), rms0, ys)
    F, Sigma = math.exp(-dt / ell), sigma ** 2 * (1 - math.exp(-2 / ell * dt))
    m, v, nell_kf = O.kalman_filter_1d(F, Sigma, 1., 1., 0.1, 0.1, ys)
    np.testing.assert_allclose(rmss[:, 1], m, rtol=1e-2)
    np.testing.assert_allclose(rmss[:, 2] - rmss[:, 1] ** 2, v, rtol=1e-3)
    np.testing.assert_allclose(nell, nell_kf, rtol=1e-5)


def test_routines_equivalence():
    """tests/test_filtering.py:113-164: rms == cms == scms filters, N=4, TME-2 (11 / 10 / 15 / 12 decimals)."""
    dt, T, ell, sigma, ys = _ou_test_data()
    b = math.sqrt(2) * sigma / math.sqrt(ell)
    N, mean0, var0 = 4, 0., 0.5
    rms0 = np.array([O.raw_moment_of_normal(mean0, var0, p) for p in range(2 * N)])
    cms0, scms0 = O.raw_to_central(rms0), O.raw_to_scaled(rms0)
    f_rms, f_cms, f_scms, f_mean, f_mean_var = O.sde_cond_moments_tme('ou', (ell,), b, dt, 2, 2 * N)
    pdf = lambda y, x: O.norm_pdf(y, x, 1.)
    rmss, nell_r = O.moment_filter_rms(f_rms, pdf, rms0, ys)
    cmss, means_c, nell_c = O.moment_filter_cms(f_cms, f_mean, pdf, cms0, mean0, ys)
    scmss, means, scales, nell_s = O.moment_filter_scms(f_scms, f_mean_var, pdf, scms0, mean0, math.sqrt(var0), ys)
    np.testing.assert_array_almost_equal(cmss, np.stack([O.raw_to_central(r) for r in rmss]), decimal=11)
    np.testing.assert_array_almost_equal(scmss, np.stack([O.raw_to_scaled(r) for r in rmss]), decimal=10)
    np.testing.assert_array_almost_equal(means_c, means, decimal=14)
    np.testing.assert_array_almost_equal(rmss[:, 2] - rmss[:, 1] ** 2, scales ** 2, decimal=12)
    for nell in (nell_s, nell_c):
        np.testing.assert_array_almost_equal(nell_r, nell, decimal=11)


@pytest.mark.parametrize('impl', ['numpy', 'c'])
def test_quadrature_integrals(impl):
    """tests/test_one_dim_quadrature.py:50-113: raw/central/scaled invariance; Gaussian-pdf / exp / sin integrands;
    polynomial exactness for Gaussian and uniform[-2, 3] measures."""
    m, v, n = 0.2, 1.1, 8
    rms = np.array([O.raw_moment_of_normal(m, v, p) for p in range(2 * n)])
    cms, sms = O.raw_to_central(rms), O.raw_to_scaled(rms)

    def quad(ms, mean=0., scale=1.):
        if impl == 'numpy':
            return O.moment_quadrature(ms, mean, scale)
        w, x = C.moment_quadrature(ms[None], mean, scale)
        return w[0], x[0]

    rules = [quad(rms), quad(cms, m), quad(sms, m, math.sqrt(v))]
    for w, x in rules:
        order = np.argsort(x)
        np.testing.assert_array_almost_equal(w[order], rules[0][0][np.argsort(rules[0][1])], decimal=6)
        np.testing.assert_array_almost_equal(x[order], np.sort(rules[0][1]), decimal=6)
        # E[exp(X)], E[sin(X)] for X ~ N(m, v)
        np.testing.assert_allclose(np.dot(w, np.exp(x)), math.exp(m + v / 2), rtol=1e-7)
        np.testing.assert_allclose(np.dot(w, np.sin(x)), math.exp(-v / 2) * math.sin(m), rtol=1e-6)
        # E[N(X; 2, 3)] = N(2; m, v + 3)   (tests/test_one_dim_quadrature.py:23-24, 62-69)
        np.testing.assert_allclose(np.dot(w, np.exp(-(x - 2) ** 2 / 6) / math.sqrt(2 * math.pi * 3)),
                                   math.exp(-(2 - m) ** 2 / (2 * (v + 3))) / math.sqrt(2 * math.pi * (v + 3)), rtol=1e-6)
        for p in range(2 * n):
            np.testing.assert_allclose(np.dot(w, x ** p), rms[p], rtol=1e-7, atol=1e-9)
    a, b = -2., 3.
    for nn in (2, 3, 4):
        ms = np.array([(b ** (p + 1) - a ** (p + 1)) / ((p + 1) * (b - a)) for p in range(2 * nn)])
        w, x = quad(ms)
        for p in range(2 * nn):
            np.testing.assert_allclose(np.dot(w, x ** p), ms[p], rtol=1e-7)


def test_tme_against_exact_ou():
    """tests/test_one_dim_moments.py:90-118: TME-3 / TME-normal vs the exact discretisation of dX = -1.1 X dt + dW."""
    dt, a, N = 0.01, -1.1, 6
    x = np.array([1.])
    F = math.exp(a * dt)
    Sig = (math.exp(2 * a * dt) - 1) / (2 * a)
    mean, variance = F * x[0], Sig
    rms = np.array([O.raw_moment_of_normal(mean, variance, p) for p in range(2 * N)])
    cms = np.array([O.central_moment_of_normal(variance, p) for p in range(2 * N)])
    c_rms, c_cms, c_scms, c_mean, c_mean_var = O.sde_cond_moments_tme('linear', (a,), 1., dt, 3, 2 * N)
    n_rms, n_cms, n_scms, n_mean, n_mean_var = O.sde_cond_moments_tme_normal('linear', (a,), 1., dt, 3, N)
    ns = np.arange(2 * N)
    np.testing.assert_allclose(mean, c_mean(x)[0], rtol=1e-7)
    np.testing.assert_allclose(variance, c_mean_var(x)[1][0], atol=1e-8, rtol=1e-6)
    np.testing.assert_allclose(mean, n_mean(x)[0], rtol=1e-7)
    np.testing.assert_allclose(variance, n_mean_var(x)[1][0], atol=1e-8, rtol=1e-6)
    np.testing.assert_allclose(rms, c_rms(x, ns)[0], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(cms, c_cms(x, ns, c_mean(x)[0])[0], atol=1e-5)
    np.testing.assert_allclose(rms, n_rms(x, ns)[0], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(cms, n_cms(x, ns, c_mean(x)[0])[0], atol=1e-8, rtol=1e-5)


def test_tme_closed_forms():
    """SURVEY.md Appendix B closed forms (sympy-derived by the surveyor, independent of this repo's derivation):
    Benes mean / variance (exact for the Benes SDE) and the 7-term raw-moment formula; well TME-2/3 mean and var."""
    dt = 1e-2
    x = np.linspace(-2.5, 2.5, 11)
    t = np.tanh(x)
    for order in (2, 3):
        mom, mean_var = O.tme_1d('benes', (), 1., dt, order, 16)
        mu, var = mean_var(x)
        np.testing.assert_allclose(mu, x + dt * t, rtol=1e-15)
        np.testing.assert_allclose(var, dt + dt ** 2 * (1 - t ** 2), rtol=1e-13)
        d3 = dt ** 3 if order == 3 else 0.
        c = [np.ones_like(x), dt * t, (dt / 2 + dt ** 2 / 2) * np.ones_like(x), (dt ** 2 / 2 + d3 / 6) * t,
             (dt ** 2 / 8 + d3 / 4) * np.ones_like(x), d3 / 8 * t, d3 / 48 * np.ones_like(x)]
        T = mom(x)
        for p in range(16):
            ref = sum(c[k] * math.perm(p, k) * x ** (p - k) for k in range(min(p, 6) + 1))
            np.testing.assert_allclose(T[:, p], ref, rtol=1e-12, atol=1e-14)
    th = 3.
    _, mv2 = O.tme_1d('well', (th,), 1., dt, 2, 4)
    _, mv3 = O.tme_1d('well', (th,), 1., dt, 3, 4)
    x = np.linspace(-1.2, 1.2, 9)
    m2 = x + dt * (x - th * x ** 3) + dt ** 2 * (3 * th ** 2 * x ** 5 / 2 - 2 * th * x ** 3 - 3 * th * x / 2 + x / 2)
    v2 = dt + dt ** 2 * (1 - 3 * th * x ** 2)
    np.testing.assert_allclose(mv2(x)[0], m2, rtol=1e-13)
    np.testing.assert_allclose(mv2(x)[1], v2, rtol=1e-13)
    m3 = m2 + dt ** 3 * (-5 * th ** 3 * x ** 7 / 2 + 9 * th ** 2 * x ** 5 / 2 + 11 * th ** 2 * x ** 3 / 2
                         - 13 * th * x ** 3 / 6 - 5 * th * x / 2 + x / 6)
    v3 = v2 + dt ** 3 * (10 * th ** 2 * x ** 4 - 8 * th * x ** 2 - 2 * th + 2 / 3)
    np.testing.assert_allclose(mv3(x)[0], m3, rtol=1e-13)
    np.testing.assert_allclose(mv3(x)[1], v3, rtol=1e-12)


def test_ldl_chol():
    """tests/test_utils.py:198-209: ldl reproduces the matrix; ldl_chol == Cholesky on a PD matrix."""
    rng = np.random.default_rng(666)
    a = rng.standard_normal((5, 5))
    mat = a @ a.T + np.eye(5)
    l, d = O.ldl(mat)
    np.testing.assert_allclose(l @ np.diag(d) @ l.T, mat, rtol=1e-12)
    np.testing.assert_allclose(O.ldl_chol(mat), np.linalg.cholesky(mat), rtol=1e-12, atol=1e-14)

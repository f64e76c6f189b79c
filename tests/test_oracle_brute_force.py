"""The NumPy oracle of the brute-force grid filter against (i) golden vectors produced by the reference's own
``mfs/classical_filters_smoothers/brute_force.py`` on the NumPy jax shim and (ii) the reference's known-answer test
(``tests/test_classical_filters_smoothers.py:204-233``: OU + Gaussian likelihood vs the Kalman filter)."""
import math
import os

import numpy as np
import pytest

from oracle import mfs_oracle as O
from oracle import mfs_oracle_bf as BF

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'golden_brute_force.npz')


def _benes_pmf(y, x):
    return O.bernoulli_pmf(y, 1 / (1 + np.exp(-x ** 3 / 5)))


def test_oracle_matches_reference_golden_benes():
    g = np.load(GOLD)
    xs, ys, ip, dt = g['benes/xs'], g['benes/ys'], g['benes/init_ps'], float(g['benes/dt'])
    keys = [k for k in g.files if k.startswith('benes/') and k.count('/') == 3]
    assert len(keys) == 15
    for key in keys:
        _, method, steps, k = key.split('/')
        out = BF.brute_force_filter('benes', (), 1., _benes_pmf, ip, xs, ys[int(k)], dt, int(steps), method)
        np.testing.assert_allclose(out, g[key], rtol=1e-12, atol=1e-14 * np.abs(g[key]).max())


def test_oracle_matches_reference_golden_ou():
    g = np.load(GOLD)
    xs, ys, ip = g['ou/xs'], g['ou/ys'], g['ou/init_ps']
    b = math.sqrt(2) * float(g['ou/sigma']) / math.sqrt(float(g['ou/ell']))
    pdf = lambda y, x: O.norm_pdf(y, x, math.sqrt(float(g['ou/r2'])))
    for key in [k for k in g.files if k.startswith('ou/') and k.count('/') == 2]:
        _, method, steps = key.split('/')
        out = BF.brute_force_filter('ou', (float(g['ou/ell']),), b, pdf, ip, xs, ys, float(g['ou/dt']), int(steps), method)
        np.testing.assert_allclose(out, g[key], rtol=1e-12, atol=1e-14 * np.abs(g[key]).max())


def ou_kalman_setting(T=40, n=1000, seed=666):
    """tests/test_classical_filters_smoothers.py:127-191 with a NumPy RNG (jax.random is not available)."""
    dt, r2 = 1e-2, 0.1
    ell, sigma = 1., 0.5
    xs = np.linspace(-5, 5, n)
    ts = np.linspace(dt, dt * T, T)
    init_ps = O.norm_pdf(xs, 0., sigma)
    F, Sigma = math.exp(-dt / ell), sigma ** 2 * (1 - math.exp(-2 * dt / ell))
    rng = np.random.Generator(np.random.PCG64(seed))
    cov = np.exp(-np.abs(ts[None, :] - ts[:, None]) / ell) * sigma ** 2
    ys = np.linalg.cholesky(cov) @ rng.standard_normal(T) + math.sqrt(r2) * rng.standard_normal(T)
    mfs_, vfs_, _ = O.kalman_filter_1d(F, Sigma, 1., r2, 0., sigma ** 2, ys)
    return dict(dt=dt, r2=r2, ell=ell, sigma=sigma, xs=xs, init_ps=init_ps, ys=ys, kf_mean=mfs_, kf_var=vfs_,
                b=math.sqrt(2) * sigma / math.sqrt(ell))


@pytest.mark.parametrize('method, atol, rtol', [('chapman-euler', 1e-4, 1e-2), ('chapman-tme-2', 1e-7, 1e-6),
                                                ('chapman-tme-3', 1e-11, 1e-10)])
def test_oracle_chapman_against_kalman(method, atol, rtol):
    s = ou_kalman_setting()
    pdf = lambda y, x: O.norm_pdf(y, x, math.sqrt(s['r2']))
    pss = BF.brute_force_filter('ou', (s['ell'],), s['b'], pdf, s['init_ps'], s['xs'], s['ys'], s['dt'], 20, method)
    m1 = BF.trapz(pss * s['xs'][None, :], s['xs'])
    m2 = BF.trapz(pss * s['xs'][None, :] ** 2, s['xs'])
    np.testing.assert_allclose(m1, s['kf_mean'], atol=atol, rtol=rtol)
    np.testing.assert_allclose(m2, s['kf_var'] + s['kf_mean'] ** 2, atol=atol, rtol=rtol)

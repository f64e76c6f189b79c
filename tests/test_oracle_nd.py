"""CPU tests of the multi-dimensional pieces: bit-exact multi-index tables (integer contract) and the nd oracle against
the golden vectors produced by the reference's own code (tests/golden/make_golden.py: golden_multi_indices, golden_nd)."""
import os

import numpy as np
import pytest

from mfs_b200.multi_dims import multi_indices as MI
from mfs_b200.multi_dims.moments import raw_moments_mvn_kan, central_moments_mvn_kan
from mfs_b200.multi_dims.ss_models import prey_predator
from oracle import mfs_oracle_nd as ND

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_multi_indices_bit_exact():
    """tests/test_multi_indices.py:8-48 (ordering, rank, cardinality) + bit-exact tables vs the reference module."""
    g = np.load(os.path.join(GOLD, 'golden_multi_indices.npz'))
    for key in g.files:
        kind, spec = key.split('/')
        nums = [int(v) for v in spec.split('_')]
        if kind == 'gen':
            d, up, lo = nums
            mine = MI.generate_graded_lexico_multi_indices(d, up, lo)
            assert mine.dtype == g[key].dtype and np.array_equal(mine, g[key])
            assert MI.sizeof_multi_indices(d, up, lo) == mine.shape[0]
            for k, row in enumerate(mine):
                assert MI.graded_lexico_indexof_multi_index(row, lo) == k
            sums = mine.sum(axis=1)
            assert np.all(np.diff(sums) >= 0) and sums.min() == lo and sums.max() == up
        else:
            N, d = nums
            mine = MI.gram_and_hankel_indices_graded_lexico(N, d)
            assert mine.dtype == g[key].dtype and np.array_equal(mine, g[key])
    # d = 2 closed-form position used by the CUDA kernel
    mis = MI.generate_graded_lexico_multi_indices(2, 11, 0)
    for k, (a, b) in enumerate(mis):
        assert (a + b) * (a + b + 1) // 2 + a == k
    assert MI.sizeof_multi_indices(3, 2, 5) == 0


def test_kan_moments_and_initial_conditions():
    g = np.load(os.path.join(GOLD, 'golden_nd.npz'))
    mean, cov = g['quad/mean'], g['quad/cov']
    for N in (3, 4, 5):
        mis = MI.generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
        ref = g[f'quad/N{N}/rms']
        kan = np.array([ND.raw_moments_mvn_kan(mean, cov, n) for n in mis])
        rec = np.array([raw_moments_mvn_kan(mean, cov, n) for n in mis])
        np.testing.assert_allclose(kan, ref, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(rec, ref, rtol=1e-11, atol=1e-13)      # recursion == Kan's sum
        np.testing.assert_allclose([central_moments_mvn_kan(cov, n) for n in mis], g[f'quad/N{N}/cms'], rtol=1e-11,
                                   atol=1e-13)
    # hand values: E[X1^2 X2] and E[X1^2 X2^2]
    m, c = np.array([0.5, -1.]), np.array([[2., 0.3], [0.3, 1.]])
    np.testing.assert_allclose(raw_moments_mvn_kan(m, c, (2, 1)), m[1] * (c[0, 0] + m[0] ** 2) + 2 * m[0] * c[0, 1])
    np.testing.assert_allclose(central_moments_mvn_kan(c, (2, 2)), c[0, 0] * c[1, 1] + 2 * c[0, 1] ** 2)
    for N in (3, 4, 5):
        mis = MI.generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
        gs = prey_predator(mis)[3]
        np.testing.assert_allclose(gs.rms, g[f'pp/N{N}/rms0'], rtol=1e-13)
        np.testing.assert_allclose(gs.cms, g[f'pp/N{N}/cms0'], rtol=1e-12, atol=1e-20)
        np.testing.assert_allclose(gs.mean, g[f'pp/N{N}/mean0'], rtol=1e-15)


def test_nd_quadrature_reproduces_moments():
    """tests/test_multi_dim_quadrature.py:71-98: all moments |n| <= 2N-1 of a Gaussian (12 decimals); nodes/weights are
    not unique (eigenvector signs), the integrated moments are."""
    g = np.load(os.path.join(GOLD, 'golden_nd.npz'))
    for N in (3, 4, 5):
        mis = MI.generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
        inds = MI.gram_and_hankel_indices_graded_lexico(N, 2)
        rms = g[f'quad/N{N}/rms']
        w, x = ND.moment_quadrature_nd(rms, inds)
        np.testing.assert_array_almost_equal(ND._powers(x, mis).T @ w, rms, decimal=11)
        np.testing.assert_array_almost_equal(ND._powers(g[f'quad/N{N}/x'], mis).T @ g[f'quad/N{N}/w'], rms, decimal=11)
        assert abs(w.sum() - 1) < 1e-12 and w.min() < 0 or N < 5          # the product rule has signed weights
        wc, xc = ND.moment_quadrature_nd(g[f'quad/N{N}/cms'], inds, g['quad/mean'])
        np.testing.assert_array_almost_equal(ND._powers(xc, mis).T @ wc, rms, decimal=11)


@pytest.mark.parametrize('N', [3])
def test_nd_oracle_filters_match_reference(N):
    g = np.load(os.path.join(GOLD, 'golden_nd.npz'))
    mis = MI.generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = MI.gram_and_hankel_indices_graded_lexico(N, 2)
    dt, T, gs = ND.prey_predator(mis)
    ys = g[f'pp/N{N}/ys']
    f_r, f_c, f_m = ND.lv_cond_moments('euler', mis)
    rmss, nell = ND.moment_filter_nd_rms(f_r, ND.lv_measurement_pmf, ys, (mis, inds), gs.rms)
    cmss, means, nell_c = ND.moment_filter_nd_cms(f_c, f_m, ND.lv_measurement_pmf, ys, (mis, inds), gs.cms, gs.mean)
    np.testing.assert_allclose(rmss, g[f'pp/N{N}/rmss'], rtol=1e-8)
    np.testing.assert_allclose(nell, g[f'pp/N{N}/nell'], rtol=1e-9)
    np.testing.assert_allclose(cmss, g[f'pp/N{N}/cmss'], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(means, g[f'pp/N{N}/means'], rtol=1e-13)
    np.testing.assert_allclose(nell_c, g[f'pp/N{N}/nell_c'], rtol=1e-13)


def test_lv_tme2_closed_forms():
    """SURVEY.md Appendix B closed forms for tme.mean_and_cov of the Lotka--Volterra SDE, order 2."""
    x = np.array([[1.1, 0.9], [0.7, 1.3], [1.0, 1.0]])
    p = ND.LV
    dt, a, b_, d_, g_, s = p['dt'], p['alp'], p['beta'], p['delta'], p['gamma'], p['sigma']
    m, c = ND.lv_mean_cov(x, 'tme_normal', 2)
    x1, x2 = x[:, 0], x[:, 1]
    m1 = x1 + dt * (a * x1 - b_ * x1 * x2) + dt ** 2 * (a ** 2 * x1 / 2 - a * b_ * x1 * x2 + b_ ** 2 * x1 * x2 ** 2 / 2
                                                       - b_ * d_ * x1 ** 2 * x2 / 2 + b_ * g_ * x1 * x2 / 2)
    m2 = x2 + dt * (d_ * x1 * x2 - g_ * x2) + dt ** 2 * (a * d_ * x1 * x2 / 2 - b_ * d_ * x1 * x2 ** 2 / 2
                                                       + d_ ** 2 * x1 ** 2 * x2 / 2 - d_ * g_ * x1 * x2 + g_ ** 2 * x2 / 2)
    np.testing.assert_allclose(m[:, 0], m1, rtol=1e-14)
    np.testing.assert_allclose(m[:, 1], m2, rtol=1e-14)
    np.testing.assert_allclose(c[:, 0, 0], s ** 2 * x1 ** 2 * dt * (1 + dt * (2 * a - 2 * b_ * x2 + s ** 2 / 2)), rtol=1e-13)
    np.testing.assert_allclose(c[:, 1, 1], s ** 2 * x2 ** 2 * dt * (1 + dt * (2 * d_ * x1 - 2 * g_ + s ** 2 / 2)), rtol=1e-13)
    np.testing.assert_allclose(c[:, 0, 1], dt ** 2 * s ** 2 * x1 * x2 * (d_ * x1 - b_ * x2) / 2, rtol=1e-12, atol=1e-20)

"""The oracle is only trusted after it is pinned: (1) against golden vectors produced by EXECUTING THE REFERENCE'S OWN
SOURCE on a NumPy jax-shim (tests/golden/make_golden.py), (2) its two restatements (NumPy/LAPACK and C) against each
other.  Tolerances are the measured rounding floor of this ill-conditioned map (Hankel Cholesky): two faithful
restatements of the same dense algorithm already differ by ~cond(G)*eps, see DESIGN.md "Parity tolerances"."""
import os

import numpy as np
import pytest

from oracle import mfs_oracle as O
from oracle import c_oracle as C
from mfs_b200.one_dim.moments import (sde_cond_moments_tme, sde_cond_moments_euler, sde_cond_moments_tme_normal)
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson
from mfs_b200.functors import linear_drift, gaussian

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def relerr(a, b):
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-300))


# ---- quadrature -----------------------------------------------------------------------------------------------------
def _quad_cases():
    g = np.load(os.path.join(GOLD, 'golden_quadrature_1d.npz'))
    names = sorted({k.split('/')[0] for k in g.files if k.split('/')[0] != 'nonpd'})
    return g, names


@pytest.mark.parametrize('impl', ['numpy', 'c'])
def test_quadrature_matches_reference(impl):
    g, names = _quad_cases()
    for name in names:
        ms, mean, scale = g[f'{name}/ms'], float(g[f'{name}/mean']), float(g[f'{name}/scale'])
        if impl == 'numpy':
            w, x = O.moment_quadrature(ms, mean, scale, sort_nodes=True)
        else:
            w, x = C.moment_quadrature(ms[None], mean, scale)
            order = np.argsort(x[0])
            w, x = w[0][order], x[0][order]
        n = len(ms) // 2
        # nodes/weights are themselves ill-conditioned functions of the moments; the reference's tests check the
        # integrals (tests/test_one_dim_quadrature.py:50-113), so do we -- plus a conditioning-scaled node check
        tol = 1e-13 * 10 ** max(0, n - 3)
        np.testing.assert_allclose(x, g[f'{name}/nodes'], rtol=tol, atol=tol)
        np.testing.assert_allclose(w, g[f'{name}/weights'], rtol=max(tol, 1e-12) * 100, atol=tol)
        assert abs(w.sum() - 1) < 1e-12


def test_quadrature_ldl_and_nonpd():
    g, names = _quad_cases()
    for name in names:
        ms, mean, scale = g[f'{name}/ms'], float(g[f'{name}/mean']), float(g[f'{name}/scale'])
        w, x = O.moment_quadrature(ms, mean, scale, sort_nodes=True, ldl=True)
        n = len(ms) // 2
        tol = 1e-12 * 10 ** max(0, n - 3)
        np.testing.assert_allclose(x, g[f'{name}/nodes_ldl'], rtol=tol, atol=tol)
    bad = g['nonpd/ms']
    w, x = O.moment_quadrature(bad)
    assert np.all(np.isnan(w)) and np.all(np.isnan(x)) and np.all(np.isnan(g['nonpd/weights']))
    wc, xc = C.moment_quadrature(bad[None])
    assert np.all(np.isnan(wc)) and np.all(np.isnan(xc))
    G = bad[np.arange(4)[:, None] + np.arange(4)[None, :]]
    np.testing.assert_allclose(O.ldl_chol(G), g['nonpd/ldl_chol'], rtol=1e-13, atol=1e-15)


def test_conversions_match_reference():
    g = np.load(os.path.join(GOLD, 'golden_conversions_1d.npz'))
    for N in (2, 5, 8):
        ic = O.gaussian_sum_1d([-0.5, 0.5], [0.05, 0.05], [0.5, 0.5], N)
        np.testing.assert_allclose(ic.rms, g[f'mix{N}/rms'], rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(ic.cms, g[f'mix{N}/cms'], rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(ic.scms, g[f'mix{N}/scms'], rtol=1e-13, atol=1e-16)
    rms = g['normal/rms']
    np.testing.assert_allclose(O.raw_to_central(rms), g['normal/raw_to_central'], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(O.raw_to_scaled(rms), g['normal/raw_to_scaled'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.central_to_raw(g['normal/raw_to_central'], 1.1), g['normal/central_to_raw'],
                               rtol=1e-12, atol=1e-9)


# ---- filters --------------------------------------------------------------------------------------------------------
# measured floor (numpy-LAPACK oracle vs C oracle vs reference-on-shim, T=100): N=5 <= 1.3e-9; N=8 raw up to 3e-5 on
# the worst of 120 trajectories (median 7e-11), central up to 1e-6.
FILTER_TOL = {5: dict(raw=5e-9, central=5e-9, scaled=2e-8, nell=1e-9), 8: dict(raw=5e-6, central=1e-7, scaled=1e-7, nell=1e-6)}


def _oracle_family(name, dt, N):
    if name.startswith('tme_normal'):
        return O.sde_cond_moments_tme_normal('benes', (), 1., dt, int(name[-1]), N)
    if name.startswith('tme'):
        return O.sde_cond_moments_tme('benes', (), 1., dt, int(name[-1]), 2 * N)
    return O.sde_cond_moments_euler('benes', (), 1., dt, N)


def _handle_family(name, dt, N, drift, disp):
    if name.startswith('tme_normal'):
        return sde_cond_moments_tme_normal(drift, disp, dt, int(name[-1]), N)
    if name.startswith('tme'):
        return sde_cond_moments_tme(drift, disp, dt, int(name[-1]))
    return sde_cond_moments_euler(drift, disp, dt, N)


@pytest.mark.parametrize('N', [5, 8])
@pytest.mark.parametrize('variant', ['euler', 'tme3'])
def test_numpy_oracle_filters_match_reference(N, variant):
    g = np.load(os.path.join(GOLD, f'golden_filter_1d_benes_N{N}.npz'))
    dt, _, ic, _, pmf = O.benes_bernoulli(N)
    np.testing.assert_array_equal(ic.rms, g['rms0'])
    fam = _oracle_family(variant, dt, N)
    tol = FILTER_TOL[N]
    for k in range(2):
        ys = g['ys'][k]
        rmss, nell = O.moment_filter_rms(fam[0], pmf, ic.rms, ys)
        assert relerr(rmss, g[f'{variant}/rms/{k}/rmss']) < tol['raw']
        assert abs(nell - float(g[f'{variant}/rms/{k}/nell'])) < tol['nell'] * abs(nell)
        cmss, means, nell = O.moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys)
        assert relerr(cmss[:, 2:], g[f'{variant}/cms/{k}/cmss'][:, 2:]) < tol['central']
        np.testing.assert_allclose(means, g[f'{variant}/cms/{k}/means'], atol=1e-10)
        assert abs(nell - float(g[f'{variant}/cms/{k}/nell'])) < tol['nell'] * abs(nell)
    if variant == 'tme3':
        scmss, means, scales, nell = O.moment_filter_scms(fam[2], fam[4], pmf, ic.scms, ic.mean,
                                                          np.sqrt(ic.variance), g['ys'][0])
        assert relerr(scmss[:, 3:], g[f'{variant}/scms/0/scmss'][:, 3:]) < tol['scaled']
        np.testing.assert_allclose(scales, g[f'{variant}/scms/0/scales'], rtol=1e-10)


@pytest.mark.parametrize('N', [5, 8])
@pytest.mark.parametrize('variant', ['euler', 'tme2', 'tme3', 'tme_normal3'])
def test_c_oracle_filters_match_reference(N, variant):
    g = np.load(os.path.join(GOLD, f'golden_filter_1d_benes_N{N}.npz'))
    dt, T, ts, ic, drift, disp, logistic, pmf, _ = benes_bernoulli(N)
    fam = _handle_family(variant, dt, N, drift, disp)
    tol = FILTER_TOL[N]
    ys = g['ys']
    o = C.filter_1d('raw', fam[0], pmf, ic.rms, ys)
    oc = C.filter_1d('central', fam[1], pmf, ic.cms, ys, mean0=ic.mean)
    for k in range(ys.shape[0]):
        assert relerr(o['ms'][k], g[f'{variant}/rms/{k}/rmss']) < tol['raw']
        assert abs(o['nell'][k] - float(g[f'{variant}/rms/{k}/nell'])) < tol['nell'] * abs(o['nell'][k])
        assert relerr(oc['ms'][k][:, 2:], g[f'{variant}/cms/{k}/cmss'][:, 2:]) < tol['central']
        np.testing.assert_allclose(oc['mean'][k], g[f'{variant}/cms/{k}/means'], atol=1e-10)
    if variant in ('tme2', 'tme3'):
        os_ = C.filter_1d('scaled', fam[2], pmf, ic.scms, ys, mean0=ic.mean, scale0=np.sqrt(ic.variance))
        for k in range(ys.shape[0]):
            assert relerr(os_['ms'][k][:, 3:], g[f'{variant}/scms/{k}/scmss'][:, 3:]) < tol['scaled']
            np.testing.assert_allclose(os_['scale'][k], g[f'{variant}/scms/{k}/scales'], rtol=1e-9)


def test_c_oracle_ou_and_well_match_reference():
    g = np.load(os.path.join(GOLD, 'golden_filter_1d_ou.npz'))
    ell, sigma, dt = float(g['ell']), float(g['sigma']), float(g['dt'])
    b = np.sqrt(2) * sigma / np.sqrt(ell)
    for N in (10, 4):
        fam = sde_cond_moments_tme(linear_drift(-1 / ell), b, dt, int(g[f'N{N}/order']))
        o = C.filter_1d('raw', fam[0], gaussian(1., 1.), g[f'N{N}/rms0'], g['ys'])
        # N=10 raw moments up to order 19 of a narrow posterior: compare the low-order, well-conditioned outputs
        np.testing.assert_allclose(o['ms'][:, 1], g[f'N{N}/rmss'][:, 1], rtol=1e-8 if N == 10 else 1e-11)
        np.testing.assert_allclose(o['ms'][:, 2], g[f'N{N}/rmss'][:, 2], rtol=1e-8 if N == 10 else 1e-11)
        assert abs(o['nell'] - float(g[f'N{N}/nell'])) < 1e-9 * abs(o['nell'])
    g = np.load(os.path.join(GOLD, 'golden_filter_1d_well.npz'))
    N = 5
    dt, T, ts, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
    th = g['theta']
    for variant, fam in (('euler', sde_cond_moments_euler(drift(th[0]), disp, dt, N)),
                         ('tme_normal2', sde_cond_moments_tme_normal(drift(th[0]), disp, dt, 2, N))):
        o = C.filter_1d('central', fam[1], pmf(th[1]), ic.cms, g['ys'], mean0=ic.mean)
        np.testing.assert_allclose(o['mean'], g[f'{variant}/means'], atol=1e-9)
        assert relerr(o['ms'][:, 2:], g[f'{variant}/cmss'][:, 2:]) < 1e-7
        assert abs(o['nell'] - float(g[f'{variant}/nell'])) < 1e-10 * abs(o['nell'])
        o = C.filter_1d('raw', fam[0], pmf(th[1]), ic.rms, g['ys'])
        assert relerr(o['ms'], g[f'{variant}/rmss']) < 1e-6


def test_stable_golden_negative_pivots():
    """golden_stable.npz (reference code on the shim): ldl=True quadratures with negative pivots and a stable=True
    filter run, NumPy oracle and C oracle."""
    from oracle import c_oracle
    g = np.load(os.path.join(GOLD, 'golden_stable.npz'))
    for name in sorted({k.split('/')[1] for k in g.files if k.startswith('quad/')}):
        w, x = O.moment_quadrature(g[f'quad/{name}/ms'], 0.3, 1.7, sort_nodes=True, ldl=True)
        np.testing.assert_allclose(x, g[f'quad/{name}/nodes'], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(w, g[f'quad/{name}/weights'], rtol=1e-10, atol=1e-15)
        wc, xc = c_oracle.moment_quadrature(g[f'quad/{name}/ms'], 0.3, 1.7, ldl=True)
        order = np.argsort(xc)
        np.testing.assert_allclose(xc[order], g[f'quad/{name}/nodes'], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(wc[order], g[f'quad/{name}/weights'], rtol=1e-8, atol=1e-12)
    fam = O.sde_cond_moments_euler('benes', (), 1., float(g['filter/dt']), 5)
    pmf = lambda y, x: O.bernoulli_pmf(y, 1 / (1 + np.exp(-x ** 3 / 5)))
    for k in range(2):
        rmss, nell = O.moment_filter_rms(fam[0], pmf, g['filter/rms0'], g['filter/ys'][k], stable=True)
        # the eps-substituted factor puts a node at ~1e8: the run is finite but chaotic in its high moments (two
        # LAPACK builds already differ in the 2nd digit there), so only the likelihood is compared, loosely
        assert np.isfinite(rmss).all() and np.isfinite(g[f'filter/{k}/rmss']).all()
        np.testing.assert_allclose(nell, g[f'filter/{k}/nell'], rtol=1e-4)


def test_characteristic_fn_golden():
    """mfs/one_dim/moments.py:309-337 (reference code on the shim) vs the oracle."""
    g = np.load(os.path.join(GOLD, 'golden_characteristic.npz'))
    zs = g['zs']
    for N in (3, 5, 8):
        for name in ('raw', 'central', 'scaled'):
            cf = np.array([O.characteristic_fn(z, g[f'mix{N}/{name}/ms'], float(g[f'mix{N}/{name}/mean']),
                                               float(g[f'mix{N}/{name}/scale'])) for z in zs])
            np.testing.assert_allclose(cf, g[f'mix{N}/{name}/cf'], rtol=0, atol=1e-11)


@pytest.mark.parametrize('N', [3, 5])
@pytest.mark.parametrize('variant', ['euler', 'tme_normal3'])
def test_c_oracle_scaled_normal_factories_match_reference(N, variant):
    """moment_filter_scms with the Normal-approximation factories: the reference divides EVERY order by
    prod(scale ** arange(2N)) (mfs/one_dim/moments.py:205, 243); the fixture is that code, executed ('euler': 100 %)."""
    g = np.load(os.path.join(GOLD, 'golden_filter_1d_scaled_normal.npz'))
    dt, T, ts, ic, drift, disp, logistic, pmf, _ = benes_bernoulli(N)
    fam = _handle_family(variant, dt, N, drift, disp)
    ys = g[f'N{N}/ys']
    o = C.filter_1d('scaled', fam[2], pmf, g[f'N{N}/scms0'], ys, mean0=float(g[f'N{N}/mean0']),
                    scale0=float(g[f'N{N}/scale0']))
    for k in range(ys.shape[0]):
        np.testing.assert_allclose(o['scale'][k], g[f'N{N}/{variant}/{k}/scales'], rtol=1e-8)
        np.testing.assert_allclose(o['mean'][k], g[f'N{N}/{variant}/{k}/means'], atol=1e-9)
        np.testing.assert_allclose(o['ms'][k][:, 2:], g[f'N{N}/{variant}/{k}/scmss'][:, 2:], rtol=1e-6, atol=1e-9)
        assert abs(o['nell'][k] - float(g[f'N{N}/{variant}/{k}/nell'])) < 1e-9 * abs(o['nell'][k])

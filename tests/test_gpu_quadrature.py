"""Batched moment_quadrature on the GPU vs the reference-generated golden vectors and the reference's own
known-answer tests (tests/test_one_dim_quadrature.py)."""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from mfs_b200.one_dim.quadtures import moment_quadrature  # noqa: E402
from mfs_b200.one_dim.moments import raw_moment_of_normal, raw_to_central, raw_to_scaled  # noqa: E402
from mfs_b200 import _lib  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_against_reference_golden():
    g = np.load(os.path.join(GOLD, 'golden_quadrature_1d.npz'))
    names = sorted({k.split('/')[0] for k in g.files if k.split('/')[0] != 'nonpd'})
    for name in names:
        ms, mean, scale = g[f'{name}/ms'], float(g[f'{name}/mean']), float(g[f'{name}/scale'])
        w, x = moment_quadrature(ms, mean, scale, sort_nodes=True)
        n = len(ms) // 2
        tol = 1e-13 * 10 ** max(0, n - 3)
        np.testing.assert_allclose(x, g[f'{name}/nodes'], rtol=tol, atol=tol)
        np.testing.assert_allclose(w, g[f'{name}/weights'], rtol=max(tol, 1e-12) * 100, atol=tol)
        assert abs(w.sum() - 1) < 1e-12
    w, x = moment_quadrature(g['nonpd/ms'])
    assert np.all(np.isnan(w)) and np.all(np.isnan(x))


def test_reference_known_answers_batched():
    m, v, n = 0.2, 1.1, 8
    rms = np.array([raw_moment_of_normal(m, v, p) for p in range(2 * n)])
    ms = torch.from_numpy(np.stack([rms, raw_to_central(rms), raw_to_scaled(rms)])).cuda()
    mean = torch.tensor([0., m, m], dtype=torch.float64, device='cuda')
    scale = torch.tensor([1., 1., math.sqrt(v)], dtype=torch.float64, device='cuda')
    w, x = moment_quadrature(ms, mean, scale, sort_nodes=True)
    assert w.is_cuda and w.shape == (3, n)
    w, x = w.cpu().numpy(), x.cpu().numpy()
    for k in range(3):
        np.testing.assert_array_almost_equal(w[k], w[0], decimal=6)
        np.testing.assert_array_almost_equal(x[k], x[0], decimal=6)
        np.testing.assert_allclose(np.dot(w[k], np.exp(x[k])), math.exp(m + v / 2), rtol=1e-7)
        np.testing.assert_allclose(np.dot(w[k], np.sin(x[k])), math.exp(-v / 2) * math.sin(m), rtol=1e-6)
        np.testing.assert_allclose(np.dot(w[k], np.exp(-(x[k] - 2) ** 2 / 6) / math.sqrt(6 * math.pi)),
                                   math.exp(-(2 - m) ** 2 / (2 * (v + 3))) / math.sqrt(2 * math.pi * (v + 3)), rtol=1e-6)
        for p in range(2 * n):
            np.testing.assert_allclose(np.dot(w[k], x[k] ** p), rms[p], rtol=1e-7, atol=1e-9)
    a, b = -2., 3.
    for nn in (1, 2, 3, 4):
        mu = np.array([(b ** (p + 1) - a ** (p + 1)) / ((p + 1) * (b - a)) for p in range(2 * nn)])
        w, x = moment_quadrature(mu)
        for p in range(2 * nn):
            np.testing.assert_allclose(np.dot(w, x ** p), mu[p], rtol=1e-7)
    # ldl=True on a PD sequence is the same rule
    w1, x1 = moment_quadrature(rms, sort_nodes=True)
    w2, x2 = moment_quadrature(rms, sort_nodes=True, ldl=True)
    np.testing.assert_array_equal(w1, w2)
    np.testing.assert_array_equal(x1, x2)


def test_quadrature_ldl_negative_pivots_against_reference_golden():
    """quadtures.py:127 with ldl=True on Hankel matrices with negative LDL pivots (dense route); weights that the
    reference itself only resolves to ~1e-17 are compared absolutely."""
    g = np.load(os.path.join(GOLD, 'golden_stable.npz'))
    names = sorted({k.split('/')[1] for k in g.files if k.startswith('quad/')})
    assert len(names) == 7
    for name in names:
        w, x = moment_quadrature(g[f'quad/{name}/ms'], 0.3, 1.7, sort_nodes=True, ldl=True)
        # the eps-substituted pivot puts |K| at ~1e8: any backward-stable eigen-solver (LAPACK in the reference, Jacobi
        # here) resolves the O(1) nodes and the weights only to ~eps |K|
        floor = 2e-14 * np.abs(g[f'quad/{name}/nodes']).max() + 1e-10
        np.testing.assert_allclose(x, g[f'quad/{name}/nodes'], rtol=1e-9, atol=floor)
        np.testing.assert_allclose(w, g[f'quad/{name}/weights'], rtol=1e-8, atol=floor)


def test_characteristic_fn_against_reference_golden():
    """characteristic_fn (mfs/one_dim/moments.py:309-337) vmapped over the z-grid and the moment history like
    dardel/benes_bernoulli/post_processing_mf.py:37-60: one launch for (T, m)."""
    from mfs_b200.one_dim.moments import characteristic_fn
    g = np.load(os.path.join(GOLD, 'golden_characteristic.npz'))
    zs = g['zs']
    for N in (3, 5, 8):
        for name in ('raw', 'central', 'scaled'):
            cf = characteristic_fn(zs, g[f'mix{N}/{name}/ms'], float(g[f'mix{N}/{name}/mean']),
                                   float(g[f'mix{N}/{name}/scale']))
            assert cf.shape == zs.shape and cf.dtype == np.complex128
            np.testing.assert_allclose(cf, g[f'mix{N}/{name}/cf'], rtol=0, atol=1e-10)
    cf = characteristic_fn(zs, torch.from_numpy(g['filter/rmss']).cuda())
    assert cf.is_cuda and cf.shape == g['filter/cf'].shape
    np.testing.assert_allclose(cf.cpu().numpy(), g['filter/cf'], rtol=0, atol=1e-9)
    one = characteristic_fn(0.7, g['mix5/raw/ms'])
    np.testing.assert_allclose(one, np.interp(0.7, zs, g['mix5/raw/cf'].real), atol=2e-2)   # scalar z -> scalar
    np.testing.assert_allclose(characteristic_fn(0., g['mix5/raw/ms']), 1. + 0.j, atol=1e-13)
    bad = g['mix5/raw/ms'].copy()
    bad[2] = bad[1] ** 2 - 1e-3
    assert np.isnan(characteristic_fn(zs, bad)).all()


def test_kernel_elementary_functions_against_libm():
    """exp / log / tanh as the kernels evaluate them (branch-free, polynomial coefficients in the constant bank) against
    NumPy's libm: a couple of ulp, over the ranges the filters use and through the special cases."""
    import ctypes
    from mfs_b200 import _lib
    rng = np.random.default_rng(9)
    x = np.concatenate([rng.uniform(-708., 708., 200000), rng.normal(size=200000) * 3., [0., -0., 1e-300, 707.9, -707.9,
                        709.9, -745., 800., -800., np.inf, -np.inf, np.nan]])
    xp = np.concatenate([np.exp(rng.uniform(-700., 700., 200000)), 1. + rng.uniform(0., 1e-3, 1000), rng.uniform(0.5, 2., 100000),
                         [1., 2.2250738585072014e-308, 1e-310, 0., -1., np.inf, np.nan, 1.7976931348623157e308]])
    L = _lib.lib()

    def run(arr, which):
        t = torch.from_numpy(arr).cuda()
        o = torch.empty_like(t)
        ptrs = [None, None, None]
        ptrs[which] = ctypes.c_void_p(o.data_ptr())
        _lib.check(L.mfs_math_selftest(t.numel(), ctypes.c_void_p(t.data_ptr()), *ptrs,
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return o.cpu().numpy()

    with np.errstate(all='ignore'):
        e, ref = run(x, 0), np.exp(x)
        fin = np.isfinite(ref) & (ref > 1e-300)
        assert np.max(np.abs(e[fin] - ref[fin]) / ref[fin]) < 4e-16
        assert np.array_equal(np.isnan(e), np.isnan(ref)) and np.array_equal(np.isinf(e), np.isinf(ref))
        assert np.all(e[~fin & ~np.isnan(ref) & ~np.isinf(ref)] < 1e-299)
        lg, ref = run(xp, 1), np.log(xp)
        fin = np.isfinite(ref)
        assert np.max(np.abs(lg[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-3)) < 1e-15
        assert np.max(np.abs(lg[fin] - ref[fin])) < 2e-13                      # |log| <= 745
        assert np.array_equal(np.isnan(lg), np.isnan(ref)) and np.array_equal(lg[~fin & ~np.isnan(ref)], ref[~fin & ~np.isnan(ref)])
        th, ref = run(x, 2), np.tanh(x)
        ok = ~np.isnan(ref)
        assert np.max(np.abs(th[ok] - ref[ok])) < 5e-16

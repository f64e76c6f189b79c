"""2-D (prey--predator) moment filter on the GPU vs the golden vectors produced by the reference's own code
(Euler--Maruyama + Kan--Magnus, mfs/multi_dims/*) and vs the NumPy nd oracle."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from mfs_b200.multi_dims.multi_indices import generate_graded_lexico_multi_indices, gram_and_hankel_indices_graded_lexico  # noqa: E402
from mfs_b200.multi_dims.filtering import moment_filter_nd_rms, moment_filter_nd_cms  # noqa: E402
from mfs_b200.multi_dims.moments import sde_cond_moments_euler_maruyama, sde_cond_moments_tme_normal  # noqa: E402
from mfs_b200.multi_dims.ss_models import prey_predator  # noqa: E402
from oracle import mfs_oracle_nd as ND  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('N', [3, 4])
def test_prey_predator_against_reference_golden(N):
    g = np.load(os.path.join(GOLD, 'golden_nd.npz'))
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, T, ts, gs, drift, dispersion, emission, pmf, _ = prey_predator(mis)
    np.testing.assert_allclose(gs.rms, g[f'pp/N{N}/rms0'], rtol=1e-13)
    np.testing.assert_allclose(gs.cms, g[f'pp/N{N}/cms0'], rtol=1e-12, atol=1e-20)
    fam = sde_cond_moments_euler_maruyama(drift, dispersion, dt, mis)
    ys = g[f'pp/N{N}/ys']
    rmss, nell = moment_filter_nd_rms((fam[0], 'index'), pmf, ys, (mis, inds), gs.rms)
    cmss, means, nell_c = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, ys, (mis, inds), gs.cms, gs.mean)
    # raw moments of a distribution concentrated at (1, 1) with std ~0.04: cond(G) is huge, the reference itself only
    # reproduces its raw-moment run to ~1e-10 between LAPACK builds; central moments are benign
    np.testing.assert_allclose(rmss, g[f'pp/N{N}/rmss'], rtol=1e-7)
    np.testing.assert_allclose(nell, g[f'pp/N{N}/nell'], rtol=1e-8)
    np.testing.assert_allclose(means, g[f'pp/N{N}/means'], rtol=1e-11)
    scale = np.abs(g[f'pp/N{N}/cmss'][:, 5:6]) ** (mis.sum(axis=1) / 2.)
    assert np.max(np.abs(cmss - g[f'pp/N{N}/cmss']) / np.maximum(np.abs(g[f'pp/N{N}/cmss']), scale)) < 1e-8
    np.testing.assert_allclose(nell_c, g[f'pp/N{N}/nell_c'], rtol=1e-11)


@pytest.mark.parametrize('family', ['euler', 'tme_normal'])
def test_paper_setting_N5_against_oracle(family):
    """BASELINE config 5 shape: N=5 (z=55, s=15, 225 nodes), central moments, batch of trajectories."""
    N, B, T = 5, 6, 12
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, simulate = prey_predator(mis)
    rng = np.random.Generator(np.random.PCG64(675))
    _, xs, ys = simulate(rng, integration_steps=10, T=T, n=B)
    fam = sde_cond_moments_euler_maruyama(drift, dispersion, dt, mis) if family == 'euler' else \
        sde_cond_moments_tme_normal(drift, dispersion, dt, 2, mis)
    cmss, means, nell, status = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, torch.from_numpy(ys).cuda(),
                                                     (mis, inds), gs.cms, gs.mean, return_status=True)
    assert cmss.shape == (B, T, 55) and means.shape == (B, T, 2)
    status = status.cpu().numpy()
    f_r, f_c, f_m = ND.lv_cond_moments(family, mis, order=2, use_kan=False)
    n_ok = 0
    for k in range(B):
        ref_c, ref_m, ref_n = ND.moment_filter_nd_cms(f_c, f_m, ND.lv_measurement_pmf, ys[k], (mis, inds), gs.cms, gs.mean)
        # unscaled central moments of a posterior with std ~0.04 up to order 9: cond(G) ~ 1e20+, so a pivot can go
        # non-positive by rounding in either implementation (the reference then returns NaN as well); compare survivors
        if not np.isfinite(ref_n) or status[k] >= 0:
            first_bad = np.argmax(~np.isfinite(ref_m[:, 0])) if not np.isfinite(ref_n) else T
            assert status[k] < 0 or abs(status[k] - first_bad) <= 3 or np.isfinite(ref_n), (k, status[k], first_bad)
            continue
        n_ok += 1
        np.testing.assert_allclose(means[k].cpu().numpy(), ref_m, rtol=1e-9)
        np.testing.assert_allclose(nell[k].item(), ref_n, rtol=1e-9)
        scale = np.abs(ref_c[:, 5:6]) ** (mis.sum(axis=1) / 2.)
        assert np.max(np.abs(cmss[k].cpu().numpy() - ref_c) / np.maximum(np.abs(ref_c), scale)) < 1e-6
    assert n_ok >= B - 2
    # history modes
    c_last, m_last, n_last = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, torch.from_numpy(ys).cuda(),
                                                  (mis, inds), gs.cms, gs.mean, history='last')
    nn = lambda t: t.nan_to_num(-7.)
    assert torch.equal(nn(c_last), nn(cmss[:, -1])) and torch.equal(nn(m_last), nn(means[:, -1])) and torch.equal(nn(n_last), nn(nell))


def test_nd_errors():
    N = 3
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, _ = prey_predator(mis)
    fam = sde_cond_moments_euler_maruyama(drift, dispersion, dt, mis)
    ys = np.zeros(4, dtype=np.uint8)
    with pytest.raises(TypeError):
        moment_filter_nd_rms((lambda x, i: x, 'index'), pmf, ys, (mis, inds), gs.rms)
    with pytest.raises(ValueError):
        moment_filter_nd_rms((fam[0], 'index'), pmf, ys, (mis[:-1], inds), gs.rms)
    with pytest.raises(ValueError):
        moment_filter_nd_rms((fam[0], 'index'), pmf, ys, (mis[::-1].copy(), inds), gs.rms)
    # non-PD initial moments -> NaN from step 0, status 0
    bad = gs.rms.copy()
    bad[3] = bad[1] ** 2 - 1e-3
    r, n, st = moment_filter_nd_rms((fam[0], 'index'), pmf, ys, (mis, inds), bad, return_status=True)
    assert np.all(np.isnan(r)) and np.isnan(n) and st == 0


@pytest.mark.parametrize('N, order', [(3, 1), (3, 2), (5, 2)])
def test_tme_without_normal_approximation_against_oracle(N, order):
    """sde_cond_moments_tme (mfs/multi_dims/moments.py:414-479, 'multi-index' signature; the paper's `tme_2` run):
    operator-coefficient evaluation on the device vs the oracle applying the generator to every monomial."""
    from mfs_b200.multi_dims.moments import sde_cond_moments_tme
    B, T = 4, 10
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, simulate = prey_predator(mis)
    rng = np.random.Generator(np.random.PCG64(676 + N))
    _, xs, ys = simulate(rng, integration_steps=10, T=T, n=B)
    fam = sde_cond_moments_tme(drift, dispersion, dt, order)
    with pytest.raises(ValueError):
        moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, ys, (mis, inds), gs.cms, gs.mean)
    cmss, means, nell, status = moment_filter_nd_cms((fam[1], 'multi-index'), fam[3], pmf, torch.from_numpy(ys).cuda(),
                                                     (mis, inds), gs.cms, gs.mean, return_status=True)
    f_r, f_c, f_m = ND.lv_cond_moments_tme(mis, order)
    status = status.cpu().numpy()
    n_ok = 0
    for k in range(B):
        ref_c, ref_m, ref_n = ND.moment_filter_nd_cms(f_c, f_m, ND.lv_measurement_pmf, ys[k], (mis, inds), gs.cms, gs.mean)
        if not np.isfinite(ref_n) or status[k] >= 0:
            continue
        n_ok += 1
        np.testing.assert_allclose(means[k].cpu().numpy(), ref_m, rtol=1e-9)
        np.testing.assert_allclose(nell[k].item(), ref_n, rtol=1e-9)
        scale = np.abs(ref_c[:, 5:6]) ** (mis.sum(axis=1) / 2.)
        assert np.max(np.abs(cmss[k].cpu().numpy() - ref_c) / np.maximum(np.abs(ref_c), scale)) < 1e-6
    assert n_ok >= B - 1
    if N == 3:   # raw moments too
        rmss, nell_r, st = moment_filter_nd_rms((fam[0], 'multi-index'), pmf, torch.from_numpy(ys).cuda(), (mis, inds),
                                                gs.rms, return_status=True)
        for k in range(B):
            ref_r, ref_n = ND.moment_filter_nd_rms(f_r, ND.lv_measurement_pmf, ys[k], (mis, inds), gs.rms)
            if np.isfinite(ref_n) and st[k].item() < 0:
                np.testing.assert_allclose(rmss[k].cpu().numpy(), ref_r, rtol=1e-6)
                np.testing.assert_allclose(nell_r[k].item(), ref_n, rtol=1e-7)


def _canon(w, x):
    """sort the nodes lexicographically (the eigenpair order of an eigh is arbitrary)"""
    order = np.lexsort((x[:, 1], x[:, 0]))
    return w[order], x[order]


def _same_measure(w, x, wr, xr, tol=1e-9):
    """K_1 / K_2 of a product-like measure have REPEATED eigenvalues, so the eigenvectors -- and with them the split of
    the weight among coincident nodes -- depend on the eigen-solver; the discrete measure sum_e w_e delta(x_e) does not.
    Compare it through a family of bounded test functions."""
    rng = np.random.Generator(np.random.PCG64(5))
    for _ in range(24):
        a, b = rng.normal(size=2), rng.uniform(0, 2 * np.pi)
        np.testing.assert_allclose(np.dot(w, np.cos(x @ a + b)), np.dot(wr, np.cos(xr @ a + b)), rtol=0, atol=tol)


def test_moment_quadrature_nd_against_reference_golden():
    """mfs/multi_dims/quadratures.py:120-178 (reference code on the shim, golden_nd.npz): correlated Gaussian, raw and
    central moments, N = 3, 4, 5; plus `ldl=True` (golden_stable.npz)."""
    from mfs_b200.multi_dims.quadratures import moment_quadrature_nd
    g = np.load(os.path.join(GOLD, 'golden_nd.npz'))
    for N in (3, 4, 5):
        inds = gram_and_hankel_indices_graded_lexico(N, 2)
        for tag, ms, mean in (('', g[f'quad/N{N}/rms'], None), ('c', g[f'quad/N{N}/cms'], g['quad/mean'])):
            w, x = moment_quadrature_nd(ms, inds, mean)
            wr, xr = g[f'quad/N{N}/w{tag}'], g[f'quad/N{N}/x{tag}']
            assert w.shape == wr.shape and x.shape == xr.shape
            (w, x), (wr, xr) = _canon(w, x), _canon(wr, xr)
            np.testing.assert_allclose(x, xr, rtol=1e-8, atol=1e-9)
            _same_measure(w, x, wr, xr)
    # batched + torch in/out
    ms = torch.from_numpy(np.stack([g['quad/N4/rms'], g['quad/N4/rms']])).cuda()
    w, x = moment_quadrature_nd(ms, gram_and_hankel_indices_graded_lexico(4, 2))
    assert w.is_cuda and w.shape == (2, 100) and x.shape == (2, 100, 2) and torch.equal(w[0], w[1])
    # ldl=True: the ordinary rule on a PD Gram matrix; with a negative pivot the eps-substituted factor puts nodes at
    # ~1e16 and nothing below eps |K| ~ 1 is resolved by ANY eigen-solver -- the large nodes and finiteness are compared
    s = np.load(os.path.join(GOLD, 'golden_stable.npz'))
    for N in (2, 3):
        inds = gram_and_hankel_indices_graded_lexico(N, 2)
        w, x = moment_quadrature_nd(s[f'nd/N{N}/ms_pd'], inds, ldl=True)
        (w, x), (wr, xr) = _canon(w, x), _canon(s[f'nd/N{N}/w_pd'], s[f'nd/N{N}/x_pd'])
        np.testing.assert_allclose(x, xr, rtol=1e-8, atol=1e-9)
        _same_measure(w, x, wr, xr)
    w, x = moment_quadrature_nd(s['nd/N2/ms'], gram_and_hankel_indices_graded_lexico(2, 2), ldl=True)
    (w, x), (wr, xr) = _canon(w, x), _canon(s['nd/N2/w'], s['nd/N2/x'])
    np.testing.assert_allclose(x, xr, rtol=1e-8, atol=1e-9)
    _same_measure(w, x, wr, xr)
    w, x = moment_quadrature_nd(s['nd/N3/ms'], gram_and_hankel_indices_graded_lexico(3, 2), ldl=True)
    assert np.isfinite(w).all() and np.isfinite(x).all()
    np.testing.assert_allclose(np.abs(x).max(), np.abs(s['nd/N3/x']).max(), rtol=1e-6)
    assert np.isnan(moment_quadrature_nd(s['nd/N3/ms'], gram_and_hankel_indices_graded_lexico(3, 2))[0]).all()


def test_paper_setting_N7_against_oracle():
    """N = 7 (z = 105 moments, s = 28 basis functions, 784 nodes): the second order the reference's prey--predator
    experiment runs (reproduce_paper_plots/plot_prey_predator_errs.py:10; on a GPU through XLA,
    dardel/run_prey_predator_mf_gpu.sh).  Round 1 stopped at N = 6."""
    N, B, T = 7, 3, 5
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, simulate = prey_predator(mis)
    rng = np.random.Generator(np.random.PCG64(679))
    _, xs, ys = simulate(rng, integration_steps=10, T=T, n=B)
    fam = sde_cond_moments_tme_normal(drift, dispersion, dt, 2, mis)
    cmss, means, nell, status = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, torch.from_numpy(ys).cuda(),
                                                     (mis, inds), gs.cms, gs.mean, return_status=True)
    assert cmss.shape == (B, T, 105) and means.shape == (B, T, 2)
    status = status.cpu().numpy()
    f_r, f_c, f_m = ND.lv_cond_moments('tme_normal', mis, order=2, use_kan=False)
    n_ok = 0
    for k in range(B):
        ref_c, ref_m, ref_n = ND.moment_filter_nd_cms(f_c, f_m, ND.lv_measurement_pmf, ys[k], (mis, inds), gs.cms, gs.mean)
        if not np.isfinite(ref_n) or status[k] >= 0:      # cond(G) ~ 1e30 at order 13: either side may lose a pivot
            continue
        n_ok += 1
        np.testing.assert_allclose(means[k].cpu().numpy(), ref_m, rtol=1e-8)
        np.testing.assert_allclose(nell[k].item(), ref_n, rtol=1e-8)
    assert n_ok >= 1
    # the quadrature on its own: moments of the initial Gaussian mixture are reproduced by the 784-node rule
    from mfs_b200.multi_dims.quadratures import moment_quadrature_nd
    w, x = moment_quadrature_nd(gs.cms, inds, mean=gs.mean)
    assert abs(w.sum() - 1.) < 1e-9
    d = x - gs.mean
    lowest = [(w * d[:, 0] ** a * d[:, 1] ** b).sum() for a, b in mis[:6]]
    np.testing.assert_allclose(lowest, gs.cms[:6], atol=1e-9)


@pytest.mark.parametrize('N', [3, 5, 6])
def test_failed_and_missing_filters_do_not_disturb_their_cta(N):
    """The 2-D kernel re-aligns the four warps (= four filters) of a CTA with barriers a few times per step.  A warp
    whose filter fails (non-PD moments -> NaN from that step on) or that has no filter at all (ragged batch) must keep
    those barriers balanced: the live filters next to it give bit-identical results to a run on their own, and nothing
    hangs."""
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, simulate = prey_predator(mis)
    fam = sde_cond_moments_tme_normal(drift, dispersion, dt, 2, mis)
    B, T = 7, 9                                                     # 7 filters = one full CTA + one with a missing warp
    rng = np.random.Generator(np.random.PCG64(31))
    _, xs, ys = simulate(rng, integration_steps=10, T=T, n=B)
    cms0 = np.tile(gs.cms, (B, 1))
    bad = [1, 4, 6]
    cms0[bad, 3] = -1e-3                                            # negative variance of x1: fails at step 0
    out = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, torch.from_numpy(ys).cuda(), (mis, inds), cms0,
                               np.tile(gs.mean, (B, 1)), return_status=True)
    cmss, means, nell, status = [o.cpu().numpy() if isinstance(o, torch.Tensor) else np.asarray(o) for o in out]
    good = [k for k in range(B) if k not in bad]
    assert np.all(status[bad] == 0) and np.all(np.isnan(nell[bad])) and np.all(np.isnan(cmss[bad]))
    assert np.all(status[good] == -1) and np.all(np.isfinite(nell[good]))
    for k in good:                                                  # each live filter alone (its CTA has 3 missing warps)
        o = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, torch.from_numpy(ys[k]).cuda(), (mis, inds), gs.cms,
                                 gs.mean)
        c1, m1, n1 = [x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x) for x in o]
        assert np.array_equal(c1, cmss[k]) and np.array_equal(m1, means[k]) and n1 == nell[k]


@pytest.mark.parametrize('mode', ['raw', 'central'])
def test_meanvar_history(mode):
    """history='meanvar' (MFS_OUT_MEANVAR for the 2-D filter): (E x1, E x2, Var x1, Cov, Var x2) per step, bit-identical
    to what the full history gives; NaN from the failing step on."""
    N, B, T = 4, 6, 11
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, simulate = prey_predator(mis)
    fam = sde_cond_moments_tme_normal(drift, dispersion, dt, 2, mis)
    rng = np.random.Generator(np.random.PCG64(5))
    _, xs, ys = simulate(rng, integration_steps=10, T=T, n=B)
    ys_t = torch.from_numpy(ys).cuda()
    if mode == 'raw':
        ms0 = np.tile(gs.rms, (B, 1))
        ms0[2, 3] = ms0[2, 1] ** 2 - 1e-3                          # record 2 fails at step 0
        full, nell = moment_filter_nd_rms((fam[0], 'index'), pmf, ys_t, (mis, inds), ms0)
        mv, nell2, st = moment_filter_nd_rms((fam[0], 'index'), pmf, ys_t, (mis, inds), ms0, history='meanvar',
                                             return_status=True)
        full = full.cpu().numpy()
        e1, e2 = full[..., 2], full[..., 1]
        want = np.stack([e1, e2, full[..., 5] - e1 * e1, full[..., 4] - e1 * e2, full[..., 3] - e2 * e2], axis=-1)
        tol = dict(rtol=1e-12, atol=1e-15)                         # fma vs separate multiply in the variance
    else:
        ms0 = np.tile(gs.cms, (B, 1))
        ms0[2, 3] = -1e-3
        full, means, nell = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, ys_t, (mis, inds), ms0,
                                                 np.tile(gs.mean, (B, 1)))
        mv, none, nell2, st = moment_filter_nd_cms((fam[1], 'index'), fam[3], pmf, ys_t, (mis, inds), ms0,
                                                   np.tile(gs.mean, (B, 1)), history='meanvar', return_status=True)
        assert none is None
        full, means = full.cpu().numpy(), means.cpu().numpy()
        want = np.stack([means[..., 0], means[..., 1], full[..., 5], full[..., 4], full[..., 3]], axis=-1)
        tol = dict(rtol=0, atol=0)
    mv, st = mv.cpu().numpy(), st.cpu().numpy()
    assert mv.shape == (B, T, 5) and st[2] == 0 and np.all(np.isnan(mv[2]))
    ok = st < 0
    np.testing.assert_allclose(mv[ok], want[ok], **tol)
    assert np.array_equal(nell.cpu().numpy(), nell2.cpu().numpy(), equal_nan=True)

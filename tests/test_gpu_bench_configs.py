"""Parity spot checks of the BASELINE configurations bench.py times at their stated sizes (configs[3], configs[4]), on a
64-record sample each: the oracle where it finishes in seconds, size-independent properties over the full length."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from mfs_b200.one_dim.filtering import moment_filter_cms  # noqa: E402
from mfs_b200.one_dim.moments import sde_cond_moments_tme_normal, sde_cond_moments_tme  # noqa: E402
from mfs_b200.one_dim.ss_models import well_poisson, benes_bernoulli  # noqa: E402
from mfs_b200.parallel import local_argmin  # noqa: E402
from mfs_b200.simulate import simulate_1d, simulate_prey_predator  # noqa: E402
from oracle import c_oracle as C  # noqa: E402


def test_config3_theta_grid_sample():
    """512 x 512 theta grid x 10^3 records (dardel/parameter_estimation/mf.py:37-73): a 6 x 6 sub-grid x 64 records, T = 1000,
    N = 7.  (i) the grid launch (one record row and one parameter row per filter, nothing materialised) equals the launch
    with every (theta, record) pair spelled out, bit for bit; (ii) nell agrees with the C oracle (the reference's dense
    algorithm) on the common survivors; (iii) so does the per-record argmin."""
    N, T, n_rec, G = 7, 1000, 64, 6
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., N)
    ys = simulate_1d(drift(3.), disp, dt, T, ic, pmf(3.), n_rec, 670)[2]
    th1, th2 = np.meshgrid(np.linspace(1.5, 4.5, G), np.linspace(1.5, 4.5, G), indexing='ij')
    th1, th2 = th1.reshape(-1), th2.reshape(-1)
    fam = sde_cond_moments_tme_normal(drift(th1[:, None]), disp, dt, 2, N)
    _, _, nell_g, st_g = moment_filter_cms(fam[1], fam[3], pmf(th2[:, None]), ic.cms, ic.mean, ys, history='none',
                                           return_status=True)
    assert nell_g.shape == (G * G, n_rec)
    ys_all = ys[None].expand(G * G, n_rec, T).contiguous()
    fam_e = sde_cond_moments_tme_normal(drift(np.broadcast_to(th1[:, None], (G * G, n_rec))), disp, dt, 2, N)
    _, _, nell_e, st_e = moment_filter_cms(fam_e[1], fam_e[3], pmf(np.broadcast_to(th2[:, None], (G * G, n_rec))), ic.cms,
                                           ic.mean, ys_all, history='none', return_status=True)
    assert torch.equal(st_g, st_e) and torch.equal(nell_g.nan_to_num(-7.), nell_e.nan_to_num(-7.))
    ref = C.filter_1d('central', fam_e[1], pmf(np.broadcast_to(th2[:, None], (G * G, n_rec))), ic.cms,
                      ys_all.cpu().numpy(), mean0=ic.mean, history='none')
    both = (st_g.cpu().numpy() < 0) & (ref['status'] < 0)
    assert both.mean() > 0.9
    rel = np.abs(nell_g.cpu().numpy()[both] - ref['nell'][both]) / np.abs(ref['nell'][both])
    assert np.median(rel) < 1e-10 and rel.max() < 1e-6, (np.median(rel), rel.max())
    _, arg = local_argmin(nell_g)
    arg_ref = np.argmin(np.where(np.isfinite(ref['nell']), ref['nell'], np.inf), axis=0)
    complete = both.all(axis=0)
    assert complete.sum() >= n_rec // 2
    assert np.array_equal(arg.cpu().numpy()[complete], arg_ref[complete])


def test_config4_prey_predator_sample():
    """Prey--predator 2-D filter, N = 5, central, tme_normal_2 and tme_2, T = 2000, 64 records
    (dardel/run_prey_predator_mf.sh:29-30).  The NumPy oracle needs ~0.3 s per step, so it checks the first 40 steps of two
    records; over the full length: rows are independent of the batch they run in, history='last' is the last row of the
    full history, central moments keep mass 1 and first moments 0, status and NaN are consistent."""
    from mfs_b200.multi_dims.multi_indices import generate_graded_lexico_multi_indices, gram_and_hankel_indices_graded_lexico
    from mfs_b200.multi_dims.filtering import moment_filter_nd_cms
    from mfs_b200.multi_dims.moments import sde_cond_moments_tme_normal as tme_normal_nd, sde_cond_moments_tme as tme_nd
    from mfs_b200.multi_dims.ss_models import prey_predator
    from oracle import mfs_oracle_nd as ND
    N, T, B = 5, 2000, 64
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt, _, ts, gs, drift, dispersion, emission, pmf, _ = prey_predator(mis)
    ys = simulate_prey_predator(drift, dispersion, dt, T, gs, pmf, B, 677)[2]
    for name, fam in (('tme_normal', tme_normal_nd(drift, dispersion, dt, 2, mis)), ('tme', tme_nd(drift, dispersion, dt, 2))):
        flag = 'index' if name == 'tme_normal' else 'multi-index'
        last, m_last, nell, st = moment_filter_nd_cms((fam[1], flag), fam[3], pmf, ys, (mis, inds), gs.cms, gs.mean,
                                                      history='last', return_status=True)
        alive = st < 0
        assert float(alive.double().mean()) > 0.5
        assert bool(torch.isfinite(nell[alive]).all()) and bool(torch.isnan(nell[~alive]).all())
        assert float((last[alive][:, 0] - 1.).abs().max()) < 1e-12
        assert float(last[alive][:, 1:3].abs().max()) < 1e-9
        sub = slice(17, 29)
        l2, m2, n2 = moment_filter_nd_cms((fam[1], flag), fam[3], pmf, ys[sub].contiguous(), (mis, inds), gs.cms, gs.mean,
                                          history='last')
        nn = lambda t: t.nan_to_num(-7.)
        assert torch.equal(nn(l2), nn(last[sub])) and torch.equal(nn(n2), nn(nell[sub])) and torch.equal(nn(m2), nn(m_last[sub]))
        full, m_full, n3 = moment_filter_nd_cms((fam[1], flag), fam[3], pmf, ys[:4, :300].contiguous(), (mis, inds), gs.cms,
                                                gs.mean)
        l4, m4, n4 = moment_filter_nd_cms((fam[1], flag), fam[3], pmf, ys[:4, :300].contiguous(), (mis, inds), gs.cms,
                                          gs.mean, history='last')
        assert torch.equal(nn(full[:, -1]), nn(l4)) and torch.equal(nn(m_full[:, -1]), nn(m4)) and torch.equal(nn(n3), nn(n4))
        if name == 'tme_normal':
            f_r, f_c, f_m = ND.lv_cond_moments('tme_normal', mis, order=2, use_kan=False)
            ys_h = ys[:2, :40].cpu().numpy()
            for k in range(2):
                ref_c, ref_m, ref_n = ND.moment_filter_nd_cms(f_c, f_m, ND.lv_measurement_pmf, ys_h[k], (mis, inds), gs.cms, gs.mean)
                np.testing.assert_allclose(m_full[k, :40].cpu().numpy(), ref_m, rtol=1e-9)


def test_config4_grid_filter_sample():
    """Brute-force grid filter at the paper's setting (dardel/benes_bernoulli/brute_force.py:21-28): n = 2000, 100
    sub-steps, chapman-tme-3, T = 100.  96 records on one grid run the sub-steps as FP64 tensor-core GEMMs; a record
    filtered alone (the reference's call shape, matrix-vector path) gives the same densities to the rounding of the
    contraction; the NumPy oracle checks the first 6 time steps of one record; mass stays 1."""
    from mfs_b200.classical_filters_smoothers import brute_force_filter
    from mfs_b200.functors import benes_drift, Dispersion, bernoulli_logistic_cubic
    from oracle import mfs_oracle as O, mfs_oracle_bf as BF
    n, steps, T, B = 2000, 100, 100, 96
    dt, _, _, ic, drift, disp, logistic, pmf, _ = benes_bernoulli(8)
    ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 678)[2]
    xs = np.linspace(-6., 6., n)
    ip = ic.pdf(xs)
    last, nell = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys, dt,
                                    integration_steps=steps, pred_method='chapman-tme-3', history='last', return_nell=True)
    mass = torch.trapezoid(last, torch.from_numpy(xs).cuda(), dim=-1)
    assert float((mass - 1.).abs().max()) < 1e-12 and bool(torch.isfinite(nell).all())
    one, nell1 = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys[37], dt,
                                    integration_steps=steps, pred_method='chapman-tme-3', history='last', return_nell=True)
    scale = float(one.abs().max())
    assert float((one - last[37]).abs().max()) / scale < 1e-11          # 10^4 chained contractions, two summation orders
    assert abs(float(nell1) - float(nell[37])) < 1e-10 * abs(float(nell1))
    full = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys[:96, :6].contiguous(),
                              dt, integration_steps=steps, pred_method='chapman-tme-3')
    _, _, _, logistic_o, pmf_o = O.benes_bernoulli(8)
    ref = BF.brute_force_filter('benes', (), 1., pmf_o, ip, xs, ys[5, :6].cpu().numpy(), dt, steps, 'chapman-tme-3')
    err = np.max(np.abs(full[5].cpu().numpy() - ref) / np.abs(ref).max(axis=-1, keepdims=True))
    assert err < 1e-11, err

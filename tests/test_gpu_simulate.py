"""Device simulators (mfs_simulate_1d / mfs_simulate_lv) against the NumPy oracle on the same Philox stream: states to
rounding, measurements equal, plus size-independent properties at full size (sharding invariance, moments of the law,
simulate -> filter pipeline)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from mfs_b200.simulate import simulate_1d, simulate_prey_predator  # noqa: E402
from mfs_b200.functors import linear_drift, gaussian  # noqa: E402
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson  # noqa: E402
from mfs_b200.one_dim.filtering import moment_filter_rms  # noqa: E402
from mfs_b200.one_dim.moments import sde_cond_moments_tme  # noqa: E402
from mfs_b200.multi_dims.ss_models import prey_predator  # noqa: E402
from mfs_b200.multi_dims import multi_indices as MI  # noqa: E402
from oracle import mfs_oracle_sim as S  # noqa: E402

IC = ([-.5, .5], [.05, .05], [.5, .5])


def _cmp(dev, ref, xtol):
    x0, xs, ys = (None if v is None else v.cpu().numpy() for v in dev)
    assert np.max(np.abs(x0 - ref[0])) < 1e-13
    assert np.max(np.abs(xs - ref[1])) < xtol, np.max(np.abs(xs - ref[1]))
    return np.mean(ys.astype(np.float64) != ref[2])


def test_benes_bernoulli_reference_simulator_vs_oracle():
    """ss_models.py:49-54: TME-3 mean_and_cov, 100 sub-steps per dt = 1e-2, T = 100 (the reference's own setting)."""
    dt, T, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(5)
    B = 300
    dev = simulate_1d(drift, disp, dt, T, ic, pmf, B, 667, return_xs=True)
    assert dev[2].dtype == torch.uint8 and dev[2].shape == (B, T)
    ref = S.simulate_1d('benes', (), 1., dt, T, *IC, 'bernoulli_logistic_cubic', (5., 0.), B, 667)
    assert _cmp(dev, ref, 1e-10) == 0.0


def test_benes_exact_law_and_odd_lengths_vs_oracle():
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(5)
    for B, T in ((257, 37), (3, 1), (129, 8)):
        dev = simulate_1d(drift, disp, dt, T, ic, pmf, B, 5, scheme='benes_exact', return_xs=True)
        ref = S.simulate_1d('benes', (), 1., dt, T, *IC, 'bernoulli_logistic_cubic', (5., 0.), B, 5, scheme='benes_exact')
        assert _cmp(dev, ref, 1e-12) == 0.0


def test_well_poisson_per_trajectory_parameters_vs_oracle():
    """dardel/parameter_estimation/mf.py:58-65 with one (theta1, theta2) pair per trajectory; orders 1..3."""
    B, T = 96, 12
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., 5)
    th1, th2 = np.linspace(0.5, 6., B), np.linspace(6., 0.5, B)
    for order, steps in ((3, 100), (2, 7), (1, 1)):
        dev = simulate_1d(drift(th1), disp, dt, T, ic, pmf(th2), B, 11, integration_steps=steps, order=order, return_xs=True)
        assert dev[2].dtype == torch.int32
        ref = [np.empty(B), np.empty((B, T)), np.empty((B, T))]
        for b in range(0, B, 32):     # the oracle takes scalar parameters: run it per trajectory group of equal theta
            for k in range(b, min(b + 32, B)):
                r = S.simulate_1d('well', (th1[k],), 1., dt, T, *IC, 'poisson_softplus', (th2[k],), 1, 11,
                                  integration_steps=steps, tme_order=order, traj_offset=k)
                ref[0][k], ref[1][k], ref[2][k] = r[0][0], r[1][0], r[2][0]
        assert _cmp(dev, ref, 1e-10) == 0.0


def test_ou_gaussian_vs_oracle_and_law():
    a, b, dt, T = -1.1, 0.7, 0.1, 5

    class IC1:
        means, variances, weights = np.array([0.4]), np.array([0.3]), np.array([1.])
    dev = simulate_1d(linear_drift(a), b, dt, T, IC1, gaussian(2., 0.5), 200, 31, integration_steps=10, return_xs=True)
    ref = S.simulate_1d('linear', (a,), b, dt, T, [0.4], [0.3], [1.], 'gaussian', (2., 0.5), 200, 31, integration_steps=10)
    x0, xs, ys = (v.cpu().numpy() for v in dev)
    assert np.max(np.abs(xs - ref[1])) < 1e-12 and np.max(np.abs(ys - ref[2])) < 1e-12
    # the law at a size the oracle would not finish: exact OU mean / variance
    B = 1 << 20
    _, xs, _ = simulate_1d(linear_drift(a), b, dt, T, IC1, gaussian(2., 0.5), B, 32, integration_steps=10, return_xs=True)
    t = dt * T
    var_exact = 0.3 * np.exp(2 * a * t) + b ** 2 / (2 * a) * (np.exp(2 * a * t) - 1)
    last = xs[:, -1]
    assert abs(float(last.mean()) - 0.4 * np.exp(a * t)) < 5 * np.sqrt(var_exact / B)
    assert abs(float(last.var()) - var_exact) < 5e-3 * var_exact


def test_sharded_batch_equals_single_batch():
    """traj_offset: two shards reproduce the single-GPU batch bit for bit (what the multi-GPU bench relies on)."""
    dt, T, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(5)
    full = simulate_1d(drift, disp, dt, 40, ic, pmf, 1000, 3, integration_steps=5, return_xs=True)
    lo = simulate_1d(drift, disp, dt, 40, ic, pmf, 600, 3, integration_steps=5, return_xs=True)
    hi = simulate_1d(drift, disp, dt, 40, ic, pmf, 400, 3, integration_steps=5, return_xs=True, traj_offset=600)
    for f, u, v in zip(full, lo, hi):
        assert torch.equal(f, torch.cat([u, v]))
    other = simulate_1d(drift, disp, dt, 40, ic, pmf, 1000, 4, integration_steps=5, return_xs=True)
    assert not torch.equal(other[1], full[1])


def test_prey_predator_milstein_vs_oracle():
    """mfs/multi_dims/ss_models.py:76-93, 100 Milstein sub-steps per dt = 1e-3."""
    mi = MI.generate_graded_lexico_multi_indices(2, 3, 0)
    dt, _, _, gs, drift, disp, _, pmf, _ = prey_predator(mi)
    B, T = 130, 40
    dev = simulate_prey_predator(drift, disp, dt, T, gs, pmf, B, 21, return_xs=True)
    ref = S.simulate_lv((4., 4., 4., 4., 0.1), dt, T, gs.means, gs.covs, gs.weights, (1., 1.), B, 21)
    x0, xs, ys = (v.cpu().numpy() for v in dev)
    assert xs.shape == (B, T, 2) and ys.shape == (B, T) and ys.dtype == np.uint8
    assert np.max(np.abs(x0 - ref[0])) < 1e-13 and np.max(np.abs(xs - ref[1])) < 1e-11
    assert np.array_equal(ys.astype(np.float64), ref[2])
    # non-diagonal initial covariance (Cholesky path of GaussianSumND.sampler)
    class GS:
        means = np.array([[1., 1.2], [0.9, 1.]])
        covs = np.array([[[2e-3, 1e-3], [1e-3, 3e-3]], [[1e-3, -5e-4], [-5e-4, 1e-3]]])
        weights = np.array([0.3, 0.7])
    dev = simulate_prey_predator(drift, disp, dt, 3, GS, pmf, 64, 22, integration_steps=3, return_xs=True, obs_dim=1)
    ref = S.simulate_lv((4., 4., 4., 4., 0.1), dt, 3, GS.means, GS.covs, GS.weights, (1., 1.), 64, 22, integration_steps=3,
                        obs_dim=1)
    assert np.max(np.abs(dev[0].cpu().numpy() - ref[0])) < 1e-13 and np.max(np.abs(dev[1].cpu().numpy() - ref[1])) < 1e-12
    assert np.array_equal(dev[2].cpu().numpy().astype(np.float64), ref[2])


def test_simulate_then_filter_pipeline_on_device():
    """The Monte-Carlo loop of dardel/benes_bernoulli/mf.py:73-92 without leaving the GPU: simulate ys, filter them."""
    N = 5
    dt, T, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    _, xs, ys = simulate_1d(drift, disp, dt, T, ic, pmf, 4096, 99, return_xs=True)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    rmss, nell, status = moment_filter_rms(fam[0], pmf, ic.rms, ys, return_status=True)
    ok = status < 0
    assert float(ok.double().mean()) > 0.95
    # the filtering mean tracks the hidden state better than the prior mean (0) does
    err_f = (rmss[ok][:, -1, 1] - xs[ok][:, -1]).pow(2).mean()
    err_0 = xs[ok][:, -1].pow(2).mean()
    assert float(err_f) < float(err_0)


def test_argument_errors():
    from mfs_b200._lib import MfsError
    dt, T, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(5)
    with pytest.raises(MfsError):
        simulate_1d(drift, disp, dt, T, ic, pmf, 8, 1, order=4)
    with pytest.raises(MfsError):
        simulate_1d(drift, disp, dt, T, ic, pmf, 8, 1, integration_steps=0)
    with pytest.raises(TypeError):
        simulate_1d(np.tanh, disp, dt, T, ic, pmf, 8, 1)
    with pytest.raises(ValueError):
        simulate_1d(drift, disp, dt, T, ic, pmf, 8, 1, device='cpu')


def test_edge_shapes():
    dt, T, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(5)
    x0, xs, ys = simulate_1d(drift, disp, dt, 5, ic, pmf, 0, 1, return_xs=True)
    assert x0.shape == (0,) and xs.shape == (0, 5) and ys.shape == (0, 5)
    x0, xs, ys = simulate_1d(drift, disp, dt, 0, ic, pmf, 9, 1, return_xs=True)
    assert x0.shape == (9,) and ys.shape == (9, 0) and bool(torch.isfinite(x0).all())
    ref = S.simulate_1d('benes', (), 1., dt, 0, *IC, 'bernoulli_logistic_cubic', (5., 0.), 9, 1)
    assert np.max(np.abs(x0.cpu().numpy() - ref[0])) < 1e-13
    # T not a multiple of 8 and an unaligned row pitch: the packed uint8 writer falls back to byte stores
    for T_ in (1, 7, 9, 15):
        _, _, y1 = simulate_1d(drift, disp, dt, T_, ic, pmf, 33, 2, scheme='benes_exact')
        r = S.simulate_1d('benes', (), 1., dt, T_, *IC, 'bernoulli_logistic_cubic', (5., 0.), 33, 2, scheme='benes_exact')
        assert np.array_equal(y1.cpu().numpy().astype(np.float64), r[2])


def test_poisson_sampler_large_rates():
    """Rates beyond the inversion range: theta2 x of a few hundred made exp(-lam) underflow and the sequential search
    run its full bound with the warp waiting (round-1 advisor finding).  Large rates take the Normal approximation with
    continuity correction; the draws equal the oracle's on the same Philox stream and have mean ~ lam, variance ~ lam."""
    B, T = 4096, 4
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., 5)
    theta2 = 800.                                        # x ~ +-0.5  ->  lam ~ 400 on the positive side (exp(theta2 x) finite)
    dev = simulate_1d(drift(3.), disp, dt, T, ic, pmf(theta2), B, 17, integration_steps=5, return_xs=True)
    ref = S.simulate_1d('well', (3.,), 1., dt, T, *IC, 'poisson_softplus', (theta2,), B, 17, integration_steps=5)
    xs, ys = dev[1].cpu().numpy(), dev[2].cpu().numpy().astype(np.float64)
    big = (theta2 * ref[1] > 100.) & (theta2 * ref[1] < 700.)
    assert big.sum() > 1000
    # the rate is theta2 * x to 1e-80 there; states agree to 1e-10, so lam to ~2e-7 and the rounded draw almost always
    assert np.mean(ys[big] == ref[2][big]) > 0.999
    lam = theta2 * xs[big]
    resid = (ys[big] - lam) / np.sqrt(lam)
    assert abs(resid.mean()) < 0.1 and abs(resid.var() - 1.) < 0.15
    assert np.all(ys[theta2 * ref[1] < -50.] == 0.)

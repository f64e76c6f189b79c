"""nell + gradient kernel (mfs_filter_1d_grad, forward-mode duals through the scan) against
  (a) the value kernel (same nell, same failure pattern),
  (b) central differences of the C oracle (the pinned restatement of the reference's dense algorithm) -- the
      reference's own gradient is jax.grad through the scan (dardel/parameter_estimation/mf.py:37-73), not runnable here,
  (c) central differences of the value kernel itself (tighter: same arithmetic on both sides).
Tolerances are those of the finite differences (h = 1e-5 relative: truncation ~1e-10, rounding ~1e-16 |nell| / h)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from mfs_b200 import synthetic  # noqa: E402
from mfs_b200.functors import linear_drift, Dispersion, gaussian, bernoulli_logistic_cubic  # noqa: E402
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms  # noqa: E402
from mfs_b200.one_dim.gradients import moment_filter_rms_value_and_grad, moment_filter_cms_value_and_grad  # noqa: E402
from mfs_b200.one_dim.moments import sde_cond_moments_tme, sde_cond_moments_tme_normal, sde_cond_moments_euler  # noqa: E402
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson  # noqa: E402
from mfs_b200.utils import GaussianSum1D  # noqa: E402
from oracle import c_oracle as C  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _fd(f, theta, k, h):
    tp, tm = list(theta), list(theta)
    tp[k], tm[k] = theta[k] * (1 + h), theta[k] * (1 - h)
    return (f(*tp) - f(*tm)) / (2 * h * theta[k])


@pytest.mark.parametrize('N', [3, 5, 7])
def test_well_poisson_central_gradient(N):
    """The reference's estimation objective: well--Poisson, central moments, TME-normal order 2 (mf.py:50-53)."""
    T = 150
    dt, _, ts, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
    ys = synthetic.well_poisson_ys_numpy(4, T, 670)
    thetas = [(3., 3.), (2.2, 3.7), (4.1, 1.9)]
    th1 = np.array([t[0] for t in thetas])[:, None]
    th2 = np.array([t[1] for t in thetas])[:, None]
    ys_b = np.broadcast_to(ys[None], (3, 4, T))
    fam = sde_cond_moments_tme_normal(drift(th1), disp, dt, 2, N)
    nell, grad, status = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(th2), ic.cms, ic.mean, dev(ys_b),
                                                          return_status=True)
    assert nell.shape == (3, 4) and grad.shape == (3, 4, 2) and bool((status < 0).all())
    _, _, nell_v = moment_filter_cms(fam[1], fam[3], pmf(th2), ic.cms, ic.mean, dev(ys_b), history='none')
    np.testing.assert_allclose(nell.cpu().numpy(), nell_v.cpu().numpy(), rtol=1e-10)

    def oracle(a, b):
        f = sde_cond_moments_tme_normal(drift(a), disp, dt, 2, N)
        return C.filter_1d('central', f[1], pmf(b), ic.cms, ys, mean0=ic.mean, history='last')['nell']

    def value_kernel(a, b):
        f = sde_cond_moments_tme_normal(drift(a), disp, dt, 2, N)
        return moment_filter_cms(f[1], f[3], pmf(b), ic.cms, ic.mean, dev(ys), history='none')[2].cpu().numpy()

    g = grad.cpu().numpy()
    for i, th in enumerate(thetas):
        for k in range(2):
            scale = np.abs(g[i, :, k]).max()
            np.testing.assert_allclose(g[i, :, k], _fd(value_kernel, th, k, 1e-5), rtol=2e-6, atol=2e-6 * scale)
            np.testing.assert_allclose(g[i, :, k], _fd(oracle, th, k, 1e-5), rtol=2e-5, atol=2e-5 * scale)


def test_raw_euler_and_single_parameter_selection():
    N, T = 4, 80
    dt, _, ts, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
    ys = synthetic.well_poisson_ys_numpy(5, T, 671)
    fam = sde_cond_moments_euler(drift(2.5), disp, dt, N)
    nell, grad = moment_filter_rms_value_and_grad(fam[0], pmf(3.2), ic.rms, dev(ys))
    only2, = moment_filter_rms_value_and_grad(fam[0], pmf(3.2), ic.rms, dev(ys), wrt=[('meas', 0)])[1:]
    assert grad.shape == (5, 2) and only2.shape == (5, 1)
    assert torch.equal(only2[:, 0], grad[:, 1])

    def value_kernel(a, b):
        f = sde_cond_moments_euler(drift(a), disp, dt, N)
        return moment_filter_rms(f[0], pmf(b), ic.rms, dev(ys), history='none')[1].cpu().numpy()

    g = grad.cpu().numpy()
    for k in range(2):
        np.testing.assert_allclose(g[:, k], _fd(value_kernel, (2.5, 3.2), k, 1e-5), rtol=2e-6,
                                   atol=2e-6 * np.abs(g[:, k]).max())


def test_ou_gaussian_three_parameters_two_passes():
    """Linear drift + Gaussian likelihood, TME-normal order 3: d nell / d (a, h, r) needs two kernel passes."""
    N, T = 4, 60
    ic = GaussianSum1D.new(np.array([0.2]), np.array([0.3]), np.array([1.]), N)
    ys = synthetic.ou_gaussian_ys_numpy(6, T, 672)
    theta = (-1.1, 0.9, 1.2)

    def fam_of(a):
        return sde_cond_moments_tme_normal(linear_drift(a), Dispersion(0.7), 0.1, 3, N)

    f = fam_of(theta[0])
    for mode in ('central', 'raw'):
        if mode == 'central':
            nell, grad = moment_filter_cms_value_and_grad(f[1], f[3], gaussian(theta[1], theta[2]), ic.cms, ic.mean, dev(ys))
            def value(a, h, r):
                fa = fam_of(a)
                return moment_filter_cms(fa[1], fa[3], gaussian(h, r), ic.cms, ic.mean, dev(ys),
                                         history='none')[2].cpu().numpy()
        else:
            nell, grad = moment_filter_rms_value_and_grad(f[0], gaussian(theta[1], theta[2]), ic.rms, dev(ys))
            value = lambda a, h, r: moment_filter_rms(fam_of(a)[0], gaussian(h, r), ic.rms, dev(ys),
                                                      history='none')[1].cpu().numpy()
        assert grad.shape == (6, 3)
        g = grad.cpu().numpy()
        np.testing.assert_allclose(nell.cpu().numpy(), value(*theta), rtol=1e-10)
        for k in range(3):
            np.testing.assert_allclose(g[:, k], _fd(value, theta, k, 1e-5), rtol=5e-6, atol=5e-6 * np.abs(g[:, k]).max())


def test_benes_tme_family_measurement_parameters():
    """TME (non-Normal) transition family, Benes drift: gradient w.r.t. the emission parameters (c0, c1).  The raw
    Benes filter is ill conditioned (nell itself is reproducible to ~1e-9 between implementations), so the finite
    differences use h = 1e-4 and carry ~1e-4 of noise."""
    N, T = 5, 100
    dt, _, ts, ic, drift, disp, logistic, pmf, _ = benes_bernoulli(N)
    ys = synthetic.benes_bernoulli_ys_numpy(8, T, 673)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    theta = (5., 0.3)
    nell, grad, st = moment_filter_rms_value_and_grad(fam[0], bernoulli_logistic_cubic(*theta), ic.rms, dev(ys),
                                                      return_status=True)
    value = lambda c0, c1: moment_filter_rms(fam[0], bernoulli_logistic_cubic(c0, c1), ic.rms, dev(ys),
                                             history='none')[1].cpu().numpy()
    ok = (st < 0).cpu().numpy()
    assert ok.mean() > 0.8
    np.testing.assert_allclose(nell.cpu().numpy()[ok], value(*theta)[ok], rtol=1e-9)
    g = grad.cpu().numpy()
    for k in range(2):
        np.testing.assert_allclose(g[ok, k], _fd(value, theta, k, 1e-4)[ok], rtol=2e-4, atol=2e-4 * np.abs(g[ok, k]).max())
    fam_c = sde_cond_moments_tme(drift, disp, dt, 2)
    nell_c, grad_c = moment_filter_cms_value_and_grad(fam_c[1], fam_c[3], bernoulli_logistic_cubic(*theta), ic.cms, ic.mean,
                                                      dev(ys))
    value_c = lambda c0, c1: moment_filter_cms(fam_c[1], fam_c[3], bernoulli_logistic_cubic(c0, c1), ic.cms, ic.mean,
                                               dev(ys), history='none')[2].cpu().numpy()
    okc = np.isfinite(nell_c.cpu().numpy())
    for k in range(2):
        np.testing.assert_allclose(grad_c.cpu().numpy()[okc, k], _fd(value_c, theta, k, 1e-4)[okc], rtol=2e-4,
                                   atol=2e-4 * np.abs(grad_c.cpu().numpy()[okc, k]).max())


def test_failure_pattern_matches_value_kernel():
    """A diverged filter gives NaN value AND gradient.  Which borderline filters lose positive definiteness depends on
    rounding (two faithful implementations differ on < 15 % of the filters at long horizons), so the
    diverged SETS are compared statistically: same fraction, mostly the same members."""
    N, B, T = 8, 2000, 300
    dt, _, ts, ic, drift, disp, logistic, pmf, _ = benes_bernoulli(N)
    ys = dev(synthetic.benes_bernoulli_ys_numpy(B, T, 674))
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    nell, grad, st = moment_filter_rms_value_and_grad(fam[0], pmf, ic.rms, ys, return_status=True)
    _, nell_v, st_v = moment_filter_rms(fam[0], pmf, ic.rms, ys, history='none', return_status=True)
    bad, bad_v = (st >= 0), (st_v >= 0)
    assert 0.02 < float(bad_v.double().mean()) < 0.9
    fb, fv, fx = float(bad.double().mean()), float(bad_v.double().mean()), float((bad != bad_v).double().mean())
    assert abs(fb - fv) < 0.03 and fx < 0.15, (fb, fv, fx)
    assert bool(torch.isnan(nell[bad]).all()) and bool(torch.isnan(grad[bad]).all())
    both = ~bad & ~bad_v
    assert bool(torch.isfinite(grad[both]).all())
    rel = ((nell[both] - nell_v[both]).abs() / nell_v[both].abs())
    assert float(rel.median()) < 1e-8      # N = 8 raw at T = 300: conditioning floor of the recursion


def test_argument_errors():
    N = 4
    dt, _, ts, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
    ys = dev(synthetic.well_poisson_ys_numpy(2, 10, 1))
    fam = sde_cond_moments_euler(drift(2.5), disp, dt, N)
    with pytest.raises(ValueError):
        moment_filter_rms_value_and_grad(fam[0], pmf(3.), ic.rms, ys, wrt=[('drift', 1)])
    with pytest.raises(ValueError):
        moment_filter_rms_value_and_grad(fam[0], pmf(3.), ic.rms, ys, wrt=[('noise', 0)])
    with pytest.raises(ValueError):
        moment_filter_rms_value_and_grad(fam[0], pmf(3.), ic.rms, ys.cpu().numpy())
    with pytest.raises(TypeError):
        moment_filter_rms_value_and_grad(np.tanh, pmf(3.), ic.rms, ys)


def test_batched_estimation_recovers_the_parameters():
    """dardel/parameter_estimation/mf.py in miniature: records simulated at theta = (3, 3), fits started at (0.1, 0.1)
    like the reference; the batched BFGS ends at stationary points and the estimates scatter around the truth."""
    from mfs_b200.simulate import simulate_1d
    from mfs_b200.one_dim.estimation import estimate_well_poisson
    N, T, B = 5, 1000, 48
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., N)
    _, _, ys = simulate_1d(drift(3.), disp, dt, T, ic, pmf(3.), B, 77, integration_steps=20)
    theta, res = estimate_well_poisson(ys, N)
    ok = res.success
    assert float(ok.double().mean()) > 0.9
    th = theta[ok]
    assert abs(float(th[:, 0].median()) - 3.) < 0.8 and abs(float(th[:, 1].median()) - 3.) < 0.5
    # stationarity in the reference's unconstrained variables
    assert float(res.grad[ok].abs().amax(1).median()) < 1e-3
    # the fit improves on the starting point for every run
    fam = sde_cond_moments_tme_normal(drift(0.1), disp, dt, 2, N)
    nell0 = moment_filter_cms(fam[1], fam[3], pmf(0.1), ic.cms, ic.mean, ys, history='none')[2]
    assert bool((res.fun[ok] < nell0[ok]).all())


def test_edge_shapes():
    """Unbatched record (T,) -> scalars like the reference's obj_func; T = 1; empty batch; non-PD start -> NaN."""
    N = 4
    dt, _, ts, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
    ys = synthetic.well_poisson_ys_numpy(3, 40, 675)
    fam = sde_cond_moments_tme_normal(drift(2.5), disp, dt, 2, N)
    nell, grad = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, dev(ys[0]))
    assert nell.shape == () and grad.shape == (2,)
    nb, gb = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, dev(ys))
    assert torch.equal(nb[0], nell) and torch.equal(gb[0], grad)
    n1, g1 = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, dev(ys[:, :1]))
    v1 = moment_filter_cms(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, dev(ys[:, :1]), history='none')[2]
    np.testing.assert_allclose(n1.cpu().numpy(), v1.cpu().numpy(), rtol=1e-12)
    assert bool(torch.isfinite(g1).all())
    n0, g0 = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(3.), ic.cms, ic.mean,
                                              torch.empty((0, 7), dtype=torch.int32, device='cuda'))
    assert n0.shape == (0,) and g0.shape == (0, 2)
    bad = ic.cms.copy()
    bad[2] = -1e-3
    nbad, gbad, sbad = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(3.), bad, ic.mean, dev(ys), return_status=True)
    assert bool(torch.isnan(nbad).all()) and bool(torch.isnan(gbad).all()) and bool((sbad == 0).all())


@pytest.mark.parametrize('N,T', [(4, 30), (6, 40)])
def test_gradient_against_complex_step_of_the_dense_algorithm(N, T):
    """Independent route to the reference's jax.grad: the complex-step derivative of a restatement of its DENSE
    algorithm in analytic arithmetic (oracle/mfs_oracle_grad.py; no finite-difference noise).  Tolerance = the
    conditioning floor of the recursion, not a differencing error."""
    from oracle import mfs_oracle_grad as G
    dt, _, ts, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
    ys = synthetic.well_poisson_ys_numpy(3, T, 680 + N)
    theta = (2.7, 3.4)
    fam = sde_cond_moments_tme_normal(drift(theta[0]), disp, dt, 2, N)
    nell, grad = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(theta[1]), ic.cms, ic.mean, dev(ys))
    ref = [G.well_poisson_value_and_grad(theta[0], theta[1], ic.cms, ic.mean, ys[k], dt=dt) for k in range(3)]
    np.testing.assert_allclose(nell.cpu().numpy(), [r[0] for r in ref], rtol=1e-10)
    g_ref = np.stack([r[1] for r in ref])
    np.testing.assert_allclose(grad.cpu().numpy(), g_ref, rtol=1e-7, atol=1e-7 * np.abs(g_ref).max())

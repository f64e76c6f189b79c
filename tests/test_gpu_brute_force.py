"""Brute-force grid filter on the GPU (FP64 tensor-core GEMM per sub-step) vs the reference-on-shim golden vectors,
the NumPy oracle and the Kalman filter (the reference's own known-answer test)."""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from mfs_b200.classical_filters_smoothers import brute_force_filter  # noqa: E402
from mfs_b200.functors import benes_drift, linear_drift, well_drift, Dispersion, bernoulli_logistic_cubic, gaussian, \
    poisson_softplus  # noqa: E402
from oracle import mfs_oracle as O  # noqa: E402
from oracle import mfs_oracle_bf as BF  # noqa: E402
from test_oracle_brute_force import ou_kalman_setting  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'golden_brute_force.npz')


def _close(out, ref, rtol=2e-13):
    """densities span hundreds of decades; compare against the row maximum (machine precision of the contraction)."""
    out = out.cpu().numpy() if hasattr(out, 'cpu') else out
    scale = np.abs(ref).max(axis=-1, keepdims=True)
    err = np.max(np.abs(out - ref) / scale)
    assert err < rtol, err


def test_benes_against_reference_golden_batched():
    g = np.load(GOLD)
    xs, ys, ip, dt = g['benes/xs'], g['benes/ys'], g['benes/init_ps'], float(g['benes/dt'])
    for method, steps in (('chapman-euler', 5), ('chapman-tme-2', 5), ('chapman-tme-3', 5), ('chapman-tme-3', 1),
                          ('kolmogorov', 20)):
        ref = np.stack([g[f'benes/{method}/{steps}/{k}'] for k in range(ys.shape[0])])
        out = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys, dt,
                                 integration_steps=steps, pred_method=method)
        assert out.shape == ref.shape
        _close(out, ref)
        # the reference's own call shape: one record (T,) -> (T, n)
        one = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys[1], dt,
                                 integration_steps=steps, pred_method=method)
        assert one.shape == ref[1].shape
        _close(one, ref[1])
        last = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys, dt,
                                  integration_steps=steps, pred_method=method, history='last')
        np.testing.assert_array_equal(last.cpu().numpy(), out[:, -1].cpu().numpy())


def test_ou_against_reference_golden():
    g = np.load(GOLD)
    ell, sigma, r2 = float(g['ou/ell']), float(g['ou/sigma']), float(g['ou/r2'])
    b = math.sqrt(2) * sigma / math.sqrt(ell)
    for method, steps in (('chapman-euler', 4), ('chapman-tme-3', 4), ('kolmogorov', 20)):
        out = brute_force_filter(linear_drift(-1. / ell), b, gaussian(1., math.sqrt(r2)), g['ou/init_ps'], g['ou/xs'],
                                 g['ou/ys'], float(g['ou/dt']), integration_steps=steps, pred_method=method)
        _close(out, g[f'ou/{method}/{steps}'])


@pytest.mark.parametrize('method, atol, rtol', [('chapman-euler', 1e-4, 1e-2), ('chapman-tme-2', 1e-7, 1e-6),
                                                ('chapman-tme-3', 1e-11, 1e-10)])
def test_chapman_against_kalman_reference_setting(method, atol, rtol):
    """tests/test_classical_filters_smoothers.py:204-225 at its full size (n=1000, T=100, 20 sub-steps)."""
    s = ou_kalman_setting(T=100)
    pss = brute_force_filter(linear_drift(-1. / s['ell']), s['b'], gaussian(1., math.sqrt(s['r2'])), s['init_ps'],
                             s['xs'], s['ys'], s['dt'], integration_steps=20, pred_method=method).cpu().numpy()
    m1 = BF.trapz(pss * s['xs'][None, :], s['xs'])
    m2 = BF.trapz(pss * s['xs'][None, :] ** 2, s['xs'])
    np.testing.assert_allclose(m1, s['kf_mean'], atol=atol, rtol=rtol)
    np.testing.assert_allclose(m2, s['kf_var'] + s['kf_mean'] ** 2, atol=atol, rtol=rtol)


def test_kolmogorov_against_kalman_reference_setting():
    """tests/test_classical_filters_smoothers.py:193-202 (rtol 1e-3 there holds for the reference's jax.random data;
    this NumPy-seeded record needs 3e-3 on one of 100 steps -- the oracle shows the same numbers, asserted below)."""
    s = ou_kalman_setting(T=100)
    pss = brute_force_filter(linear_drift(-1. / s['ell']), s['b'], gaussian(1., math.sqrt(s['r2'])), s['init_ps'],
                             s['xs'], s['ys'], s['dt'], integration_steps=20, pred_method='kolmogorov').cpu().numpy()
    m1 = BF.trapz(pss * s['xs'][None, :], s['xs'])
    m2 = BF.trapz(pss * s['xs'][None, :] ** 2, s['xs'])
    np.testing.assert_allclose(m1, s['kf_mean'], atol=1e-4, rtol=1e-2)
    np.testing.assert_allclose(m2, s['kf_var'] + s['kf_mean'] ** 2, atol=1e-4, rtol=3e-3)
    ref = BF.brute_force_filter('ou', (s['ell'],), s['b'], lambda y, x: O.norm_pdf(y, x, math.sqrt(s['r2'])),
                                s['init_ps'], s['xs'], s['ys'][:10], s['dt'], 20, 'kolmogorov')
    _close(pss[:10], ref, rtol=1e-12)


@pytest.mark.parametrize('n, B', [(333, 5), (250, 131), (1024, 2)])
def test_ragged_sizes_against_oracle(n, B):
    """grid sizes that are odd / not multiples of the 16-wide contraction slice or the 128-wide tiles; batch sizes that
    are not multiples of the 128-row tile; per-record initial densities and measurement parameters."""
    rng = np.random.Generator(np.random.PCG64(680 + n))
    T, steps, dt = 4, 3, 1e-2
    xs = np.linspace(-3.5, 3.7, n)
    ys = rng.poisson(2., size=(B, T)).astype(np.int32)
    mus = rng.uniform(-0.5, 0.5, B)
    ip = O.norm_pdf(xs[None, :], mus[:, None], 0.4)
    theta2 = rng.uniform(1., 3., B)
    out, nell = brute_force_filter(well_drift(2.), Dispersion(1.), poisson_softplus(theta2), ip, xs, ys, dt,
                                   integration_steps=steps, pred_method='chapman-tme-2', return_nell=True)
    for k in (0, B // 2, B - 1):
        pmf = lambda y, x: O.poisson_pmf(y, np.log1p(np.exp(theta2[k] * x)))
        ref = BF.brute_force_filter('well', (2.,), 1., pmf, ip[k], xs, ys[k], dt, steps, 'chapman-tme-2')
        _close(out[k], ref)
    mass = BF.trapz(out.cpu().numpy(), xs)
    np.testing.assert_allclose(mass, 1., rtol=1e-13)
    assert np.isfinite(nell.cpu().numpy()).all()


def test_paper_grid_size_properties():
    """n = 2000 (dardel/benes_bernoulli/brute_force.py:22), 300 records: unit mass after every update; batched (GEMM)
    and record-by-record (matrix-vector) runs agree."""
    rng = np.random.Generator(np.random.PCG64(681))
    n, B, T, steps = 2000, 300, 3, 4
    xs = np.linspace(-6., 6., n)
    ys = (rng.random((B, T)) < 0.5).astype(np.uint8)
    ip = 0.5 * O.norm_pdf(xs, -0.5, math.sqrt(0.05)) + 0.5 * O.norm_pdf(xs, 0.5, math.sqrt(0.05))
    args = (benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs)
    out = brute_force_filter(*args, ys, 1e-2, integration_steps=steps, pred_method='chapman-tme-3')
    np.testing.assert_allclose(BF.trapz(out.cpu().numpy(), xs), 1., rtol=1e-13)
    # rows of the contraction are independent: a different batch composition gives the same bits (GEMM path) ...
    sub = brute_force_filter(*args, ys[100:300], 1e-2, integration_steps=steps, pred_method='chapman-tme-3')
    np.testing.assert_array_equal(sub.cpu().numpy(), out[100:300].cpu().numpy())
    # ... and the small-batch matrix-vector path (one record, 5 records) agrees to the summation order
    for k in (0, 127, 128, 299):
        one = brute_force_filter(*args, ys[k], 1e-2, integration_steps=steps, pred_method='chapman-tme-3')
        _close(one, out[k].cpu().numpy(), rtol=1e-13)
    five = brute_force_filter(*args, ys[10:15], 1e-2, integration_steps=steps, pred_method='chapman-tme-3')
    _close(five, out[10:15].cpu().numpy(), rtol=1e-13)
    ref = BF.brute_force_filter('benes', (), 1., lambda y, x: O.bernoulli_pmf(y, 1 / (1 + np.exp(-x ** 3 / 5))), ip, xs,
                                ys[7], 1e-2, steps, 'chapman-tme-3')
    _close(out[7], ref)


def test_rejects_python_callables_and_bad_methods():
    xs = np.linspace(-1, 1, 16)
    with pytest.raises(TypeError):
        brute_force_filter(np.tanh, 1., bernoulli_logistic_cubic(), xs, xs, np.zeros(3, np.uint8), 1e-2)
    with pytest.raises(NotImplementedError):
        brute_force_filter(benes_drift(), 1., bernoulli_logistic_cubic(), xs, xs, np.zeros(3, np.uint8), 1e-2,
                           pred_method='magic')


def test_per_record_grids_like_the_reference_experiment():
    """dardel/benes_bernoulli/brute_force.py:73-76: every record is filtered on its OWN grid (spanned by its moment-filter
    run).  xs (B, n): equals the single-record calls bit for bit and the NumPy oracle on each record's grid."""
    rng = np.random.default_rng(5)
    B, T, n = 5, 8, 300
    dt = 1e-2
    _, _, _, logistic, pmf_o = O.benes_bernoulli(5)
    from mfs_b200.one_dim.ss_models import benes_bernoulli
    ic = benes_bernoulli(5)[3]
    ys = (rng.random((B, T)) < 0.5).astype(np.uint8)
    lb, ub = -3. - rng.random(B), 3. + rng.random(B)
    xs = np.stack([np.linspace(lb[k], ub[k], n) for k in range(B)])
    ip = np.stack([ic.pdf(xs[k]) for k in range(B)])
    out, nell = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys, dt,
                                   integration_steps=7, pred_method='chapman-tme-3', return_nell=True)
    assert out.shape == (B, T, n) and nell.shape == (B,)
    for k in range(B):
        one = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip[k], xs[k], ys[k], dt,
                                 integration_steps=7, pred_method='chapman-tme-3')
        assert torch.equal(one, out[k])
        ref = BF.brute_force_filter('benes', (), 1., pmf_o, ip[k], xs[k], ys[k], dt, 7, 'chapman-tme-3')
        _close(out[k], ref)
    with pytest.raises(ValueError):
        brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip[0], xs, ys, dt)


@pytest.mark.parametrize('steps', [2, 5, 12, 100])
def test_power_operator_equals_the_literal_recursion(steps):
    """power_operator=True: the k-th power of the sub-step operator by binary powering (k = 2 / 5 / 12 / 100: even, odd,
    mixed bit patterns), one contraction per time step.  Equal to the literal k-sub-step recursion up to the order of
    summation (all terms non-negative), on the GEMM path, the matrix-vector path and per-record grids; and to the NumPy
    oracle of the reference's recursion."""
    rng = np.random.Generator(np.random.PCG64(90 + steps))
    n, B, T = 300, 150, 4
    xs = np.linspace(-5., 5., n)
    ys = (rng.random((B, T)) < 0.5).astype(np.uint8)
    ip = 0.5 * O.norm_pdf(xs, -0.5, math.sqrt(0.05)) + 0.5 * O.norm_pdf(xs, 0.5, math.sqrt(0.05))
    args = (benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs)
    kw = dict(integration_steps=steps, pred_method='chapman-tme-3', return_nell=True)
    lit, nell_l = brute_force_filter(*args, ys, 1e-2, **kw)
    pw, nell_p = brute_force_filter(*args, ys, 1e-2, power_operator=True, **kw)
    _close(pw, lit.cpu().numpy(), rtol=1e-11)
    np.testing.assert_allclose(nell_p.cpu().numpy(), nell_l.cpu().numpy(), rtol=1e-11)
    one = brute_force_filter(*args, ys[3], 1e-2, power_operator=True, integration_steps=steps, pred_method='chapman-tme-3')
    _close(one, lit[3].cpu().numpy(), rtol=1e-11)                      # matrix-vector path on the powered operator
    ref = BF.brute_force_filter('benes', (), 1., lambda y, x: O.bernoulli_pmf(y, 1 / (1 + np.exp(-x ** 3 / 5))), ip, xs,
                                ys[3], 1e-2, steps, 'chapman-tme-3')
    _close(pw[3], ref, rtol=1e-11)
    if steps == 5:                                                     # per-record grids, and 'kolmogorov' ignores the flag
        xs2 = np.stack([np.linspace(-5. - 0.1 * k, 5. + 0.1 * k, n) for k in range(3)])
        ip2 = np.stack([0.5 * O.norm_pdf(x, -0.5, math.sqrt(0.05)) + 0.5 * O.norm_pdf(x, 0.5, math.sqrt(0.05)) for x in xs2])
        a2 = (benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip2, xs2, ys[:3], 1e-2)
        _close(brute_force_filter(*a2, integration_steps=steps, pred_method='chapman-euler', power_operator=True),
               brute_force_filter(*a2, integration_steps=steps, pred_method='chapman-euler').cpu().numpy(), rtol=1e-11)
        k1 = brute_force_filter(*args, ys[:2], 1e-3, integration_steps=20, pred_method='kolmogorov', power_operator=True)
        k0 = brute_force_filter(*args, ys[:2], 1e-3, integration_steps=20, pred_method='kolmogorov')
        assert torch.equal(k1, k0)

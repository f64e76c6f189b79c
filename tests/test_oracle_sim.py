"""CPU tests of the simulator oracle (oracle/mfs_oracle_sim.py): the Philox4x32-10 restatement against the published
known-answer vectors of the Random123 distribution (kat_vectors: philox4x32 10), the uniform / Box--Muller maps, and the
law of the simulators against closed forms (exact OU transition; exact Benes moments E[X_t] from the reference's
model)."""
import numpy as np

from oracle import mfs_oracle_sim as S


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, ref in kat:
        out = S.philox4x32_10(np.array(ctr, dtype=np.uint64), key)
        assert [int(v) for v in out] == list(ref)
    batch = S.philox4x32_10(np.array([k[0] for k in kat[:1] * 3], dtype=np.uint64), (0, 0))
    assert batch.shape == (3, 4) and np.all(batch == batch[0])


def test_uniforms_and_normals():
    ua, ub = S.draw(np.arange(400000), 3, 7, 12345)
    assert ua.min() > 0 and ua.max() < 1 and ub.min() > 0 and ub.max() < 1
    assert abs(ua.mean() - 0.5) < 2e-3 and abs(ub.var() - 1 / 12) < 1e-3
    assert abs(np.corrcoef(ua, ub)[0, 1]) < 5e-3
    z = np.concatenate(S.normals(ua, ub))
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3 and abs((z ** 4).mean() - 3) < 3e-2
    # different (time, draw, seed, trajectory) -> different numbers
    assert not np.any(S.draw(np.arange(8), 3, 7, 12345)[0] == S.draw(np.arange(8), 3, 8, 12345)[0])
    assert not np.any(S.draw(np.arange(8), 3, 7, 12345)[0] == S.draw(np.arange(8), 4, 7, 12345)[0])
    assert not np.any(S.draw(np.arange(8), 3, 7, 12345)[0] == S.draw(np.arange(8), 3, 7, 12346)[0])
    assert np.array_equal(S.draw(np.arange(8) + 2 ** 33, 3, 7, 1)[0][:4], S.draw(np.arange(4) + 2 ** 33, 3, 7, 1)[0])


def test_sharded_stream_equals_single_batch():
    kw = dict(integration_steps=4, tme_order=3)
    full = S.simulate_1d('well', (3.,), 1., 1e-2, 6, [-.5, .5], [.05, .05], [.5, .5], 'poisson_softplus', (3.,), 40, 9, **kw)
    lo = S.simulate_1d('well', (3.,), 1., 1e-2, 6, [-.5, .5], [.05, .05], [.5, .5], 'poisson_softplus', (3.,), 25, 9, **kw)
    hi = S.simulate_1d('well', (3.,), 1., 1e-2, 6, [-.5, .5], [.05, .05], [.5, .5], 'poisson_softplus', (3.,), 15, 9,
                       traj_offset=25, **kw)
    for f, a, b in zip(full, lo, hi):
        assert np.array_equal(f, np.concatenate([a, b]))


def test_ou_law():
    """simulate_sde on dX = a X dt + b dW with TME-3 sub-steps reproduces the exact OU transition moments."""
    a, b, dt, T, B = -1.1, 0.7, 0.1, 5, 20000
    x0, xs, ys = S.simulate_1d('linear', (a,), b, dt, T, [0.4], [0.3], [1.], 'gaussian', (2., 0.5), B, 31,
                               integration_steps=10)
    t = dt * T
    mean_exact = 0.4 * np.exp(a * t)
    var_exact = 0.3 * np.exp(2 * a * t) + b ** 2 / (2 * a) * (np.exp(2 * a * t) - 1)
    se = np.sqrt(var_exact / B)
    assert abs(xs[:, -1].mean() - mean_exact) < 4 * se
    assert abs(xs[:, -1].var() - var_exact) < 0.03 * var_exact
    assert abs(x0.mean() - 0.4) < 4 * np.sqrt(0.3 / B) and abs(x0.var() - 0.3) < 0.01
    resid = ys[:, -1] - 2. * xs[:, -1]
    assert abs(resid.std() - 0.5) < 0.01 and abs(resid.mean()) < 0.02


def test_benes_tme_substeps_match_exact_law():
    """The reference's simulator (100 TME-3 Gaussian sub-steps, ss_models.py:49-54) and the exact Benes transition
    agree in distribution (mean, variance, Bernoulli rate), and both match the closed form of the exact law
    X_t | x0 ~ sum_+- (1 +- tanh x0)/2 N(x0 +- t, t):  E[X_t^2 | x0] = x0^2 + t + t^2 + 2 t x0 tanh x0."""
    kw = dict(B=20000, seed=5)
    ic = ([-.5, .5], [.05, .05], [.5, .5])
    _, xa, ya = S.simulate_1d('benes', (), 1., 1e-2, 30, *ic, 'bernoulli_logistic_cubic', (5., 0.), integration_steps=4, **kw)
    _, xb, yb = S.simulate_1d('benes', (), 1., 1e-2, 30, *ic, 'bernoulli_logistic_cubic', (5., 0.), scheme='benes_exact',
                              B=20000, seed=6)
    a, b = xa[:, -1], xb[:, -1]
    assert abs(a.mean() - b.mean()) < 0.03
    assert abs(a.var() - b.var()) < 0.03 * b.var() + 0.01
    t = 0.3
    x0 = S.simulate_1d('benes', (), 1., 1e-2, 0, *ic, 'bernoulli_logistic_cubic', (5., 0.), **kw)[0]
    m2 = np.mean(x0 ** 2 + t + 2 * t * x0 * np.tanh(x0) + t ** 2)
    assert abs((b ** 2).mean() - m2) < 0.03 and abs((a ** 2).mean() - m2) < 0.03
    assert abs(ya.mean() - yb.mean()) < 0.01


def test_lv_milstein_small_noise_follows_the_ode():
    """sigma -> 0: the Milstein scheme reduces to explicit Euler on the Lotka--Volterra ODE, whose invariant
    V = delta x - gamma log x + beta y - alpha log y is conserved up to O(dt)."""
    p = (4., 4., 4., 4., 1e-9)
    means = np.array([[1.2, 0.8]])
    covs = np.array([np.eye(2) * 1e-12])
    x0, xs, ys = S.simulate_lv(p, 1e-3, 50, means, covs, [1.], (1., 1.), 4, 3, integration_steps=20)
    V = lambda x: 4 * x[..., 0] - 4 * np.log(x[..., 0]) + 4 * x[..., 1] - 4 * np.log(x[..., 1])
    assert np.max(np.abs(V(xs) - V(x0)[:, None])) < 1e-4
    assert set(np.unique(ys)) <= {0., 1.}

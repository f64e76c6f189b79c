"""CPU tests of the gradient oracle (oracle/mfs_oracle_grad.py): its VALUE equals the C oracle's nell (the pinned
restatement of the reference's dense algorithm) to rounding, and its complex-step DERIVATIVE equals central differences
of that oracle to their noise level -- for the reference's estimation objective (dardel/parameter_estimation/mf.py:37-54:
well--Poisson, central moments, TME-normal order 2) and its --euler variant."""
import numpy as np
import pytest

from mfs_b200.one_dim.moments import sde_cond_moments_tme_normal, sde_cond_moments_euler
from mfs_b200.one_dim.ss_models import well_poisson
from mfs_b200.synthetic import well_poisson_ys_numpy
from oracle import c_oracle as C
from oracle import mfs_oracle as O
from oracle import mfs_oracle_grad as G


@pytest.mark.parametrize('N,T,theta,order', [(4, 30, (2.5, 3.2), 2), (5, 40, (4.1, 1.9), 2), (4, 30, (3.0, 3.0), 1)])
def test_complex_step_oracle_matches_the_dense_oracle(N, T, theta, order):
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., N)
    ys = well_poisson_ys_numpy(2, T, 7)

    def dense_nell(a, b):
        f = sde_cond_moments_tme_normal(drift(a), disp, dt, 2, N) if order == 2 else sde_cond_moments_euler(drift(a), disp, dt, N)
        return C.filter_1d('central', f[1], pmf(b), ic.cms, ys, mean0=ic.mean, history='last')['nell']

    res = [G.well_poisson_value_and_grad(theta[0], theta[1], ic.cms, ic.mean, ys[k], dt=dt, order=order) for k in range(2)]
    val, grad = np.array([r[0] for r in res]), np.stack([r[1] for r in res])
    np.testing.assert_allclose(val, dense_nell(*theta), rtol=1e-12)
    h = 1e-5
    fd = np.stack([(dense_nell(theta[0] * (1 + h), theta[1]) - dense_nell(theta[0] * (1 - h), theta[1])) / (2 * h * theta[0]),
                   (dense_nell(theta[0], theta[1] * (1 + h)) - dense_nell(theta[0], theta[1] * (1 - h))) / (2 * h * theta[1])], axis=1)
    np.testing.assert_allclose(grad, fd, rtol=1e-7, atol=1e-7 * np.abs(fd).max())


def test_pieces_against_the_numpy_oracle():
    """The analytic building blocks at real arguments: quadrature = LAPACK route, TME mean / variance = definition-driven
    TME of the NumPy oracle, Normal moments = the reference's binomial sum."""
    rng = np.random.default_rng(3)
    x, w = np.sort(rng.normal(size=5)), rng.random(5) + 0.1
    w /= w.sum()
    ms = np.array([np.sum(w * x ** p) for p in range(10)])
    wq, xq = G.moment_quadrature(ms.astype(complex), 0.3 + 0j)
    wo, xo = O.moment_quadrature(ms, 0.3)
    iq, io = np.argsort(xq.real), np.argsort(xo)
    np.testing.assert_allclose(xq.real[iq], xo[io], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(wq.real[iq], wo[io], rtol=1e-7, atol=1e-11)
    assert np.abs(xq.imag).max() == 0 and np.abs(wq.imag).max() == 0
    _, mean_var = O.tme_1d('well', (2.7,), 1., 1e-2, 2, 2)
    mu, var = G.well_mean_var(x, 2.7, 1e-2)
    np.testing.assert_allclose(mu, mean_var(x)[0], rtol=1e-14)
    np.testing.assert_allclose(var, mean_var(x)[1], rtol=1e-14)
    mom = G._normal_moments(0.4, 0.09, 8)
    np.testing.assert_allclose(mom, [O.raw_moment_of_normal(0.4, 0.09, p) for p in range(8)], rtol=1e-14)
    # a Gram matrix that is not positive definite gives NaN, like the scan
    assert np.isnan(G.well_poisson_nell(3., 3., np.array([1., 0., -1e-3, 0.]), 0., [1, 0]).real)

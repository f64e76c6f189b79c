"""Host-side mirror of ``mfs/multi_dims/ss_models.py``: the prey--predator (Lotka--Volterra) model."""
import numpy as np

from ..functors import lotka_volterra_drift, proportional_dispersion, bernoulli_logistic_cubic
from ..utils import GaussianSumND

__all__ = ['prey_predator']


def prey_predator(multi_indices):
    """``mfs/multi_dims/ss_models.py:40-95``.  Returns ``dt, T, ts, gs, drift, dispersion, emission,
    measurement_cond_pmf, simulate`` like the reference; drift / dispersion / pmf are functor handles (the Bernoulli
    observation acts on x[0] with emission 1/(1+exp(-x^3+1))); ``simulate(rng, integration_steps)`` is the Milstein
    scheme of ``:69-93`` on a NumPy generator (data generation only)."""
    dt = 1e-3
    T = 2000
    ts = np.linspace(dt, dt * T, T)
    alp, beta, delta, gamma, sigma = 4., 4., 4., 4., 0.1
    means = np.array([[1., 1.], [1., 1.]])
    covs = np.array([[[1., 0.], [0., 1.]], [[2., 0.], [0., 2.]]]) * 0.001
    weights = np.array([0.5, 0.5])
    gs = GaussianSumND.new(means, covs, weights, multi_indices)
    drift = lotka_volterra_drift(alp, beta, delta, gamma)
    dispersion = proportional_dispersion(sigma)

    def emission(x):
        with np.errstate(over='ignore'):
            return 1 / (1 + np.exp(-np.asarray(x, dtype=np.float64) ** 3 + 1))

    measurement_cond_pmf = bernoulli_logistic_cubic(1., 1.)

    def simulate(rng: np.random.Generator, integration_steps: int = 100, T: int = T, n: int = 1):
        x = gs.sampler(rng, n)
        ddt = dt / integration_steps
        xs = np.empty((T, n, 2))
        for t in range(T):
            for _ in range(integration_steps):
                ddw = np.sqrt(ddt) * rng.standard_normal((n, 2))
                a = x * (x[:, ::-1] * np.array([-beta, delta]) + np.array([alp, -gamma]))
                x = x + a * ddt + sigma * x * ddw + 0.5 * sigma ** 2 * x * (ddw ** 2 - ddt)
            xs[t] = x
        ys = (rng.random((T, n)) < emission(xs[:, :, 0])).astype(np.uint8)
        return x, np.swapaxes(xs, 0, 1), ys.T.copy()

    return dt, T, ts, gs, drift, dispersion, emission, measurement_cond_pmf, simulate

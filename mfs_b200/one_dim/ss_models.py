"""Host-side mirror of ``mfs/one_dim/ss_models.py``: model constants, functor handles and NumPy data simulators."""
import numpy as np

from ..functors import benes_drift, well_drift, Dispersion, bernoulli_logistic_cubic, poisson_softplus
from ..utils import GaussianSum1D

__all__ = ['benes_bernoulli', 'well_poisson']


def _init_cond(N):
    return GaussianSum1D.new(means=np.array([-0.5, 0.5]), variances=np.array([0.05, 0.05]),
                             weights=np.array([0.5, 0.5]), N=N)


def benes_bernoulli(N: int = 2):
    """The Benes--Bernoulli model (``mfs/one_dim/ss_models.py:25-56``).

    Returns ``dt, T, ts, init_cond, drift, dispersion, logistic, measurement_cond_pmf, simulate_trajectory`` in the
    reference's order; ``drift`` / ``dispersion`` / ``measurement_cond_pmf`` are functor handles.
    ``simulate_trajectory(x0, rng, T=T)`` draws from the *exact* Benes transition law (mixture of N(x +- dt, dt) with
    probabilities (1 +- tanh x)/2) instead of the reference's 100 TME-3 Gaussian sub-steps: data generation only.
    """
    dt = 1e-2
    T = 100
    ts = np.linspace(dt, dt * T, T)
    init_cond = _init_cond(N)
    drift = benes_drift()
    dispersion = Dispersion(1.)

    def logistic(x):
        return 1 / (1 + np.exp(-np.asarray(x, dtype=np.float64) ** 3 / 5))

    measurement_cond_pmf = bernoulli_logistic_cubic(5., 0.)

    def simulate_trajectory(x0, rng: np.random.Generator, T: int = T):
        x = np.array(x0, dtype=np.float64, copy=True)
        xs = np.empty((T,) + x.shape)
        for t in range(T):
            s = np.where(rng.random(x.shape) < 0.5 * (1 + np.tanh(x)), 1., -1.)
            x = x + s * dt + np.sqrt(dt) * rng.standard_normal(x.shape)
            xs[t] = x
        return xs

    return dt, T, ts, init_cond, drift, dispersion, logistic, measurement_cond_pmf, simulate_trajectory


def well_poisson(true_p1, N: int = 2):
    """The Well--Poisson model (``mfs/one_dim/ss_models.py:59-93``).  ``drift`` and ``measurement_cond_pmf`` are
    *factories* of handles taking the parameter, like the reference's two-argument callables
    ``drift(x, p)`` / ``measurement_cond_pmf(y, x, p)``."""
    dt = 1e-2
    T = 1000
    ts = np.linspace(dt, dt * T, T)
    init_cond = _init_cond(N)

    def drift(p):
        return well_drift(p)

    dispersion = Dispersion(1.)

    def emission(x, p):
        return np.log(1. + np.exp(p * np.asarray(x, dtype=np.float64)))

    def measurement_cond_pmf(p):
        return poisson_softplus(p)

    def simulate_trajectory(x0, rng: np.random.Generator, T: int = T, integration_steps: int = 100):
        x = np.array(x0, dtype=np.float64, copy=True)
        xs = np.empty((T,) + x.shape)
        h = dt / integration_steps
        for t in range(T):
            for _ in range(integration_steps):
                x = x + x * (1 - true_p1 * x ** 2) * h + np.sqrt(h) * rng.standard_normal(x.shape)
            xs[t] = x
        return xs

    return dt, T, ts, init_cond, drift, dispersion, emission, measurement_cond_pmf, simulate_trajectory

"""Seeded synthetic workloads of the BASELINE configs (data generation only -- not part of the filter path).

The reference draws its data with ``jax.random`` keys from ``rng_keys.npy`` (``dardel/benes_bernoulli/mf.py:73-80``),
which cannot be reproduced without JAX; these generators define our own seeded equivalents with the same shapes and
dtypes.  Benes trajectories use the exact transition law (mixture of N(x +- dt, dt), P(+-) = (1 +- tanh x)/2) instead
of the reference's 100 TME-3 Gaussian sub-steps (``mfs/one_dim/ss_models.py:50-54``).
"""
import numpy as np


def benes_bernoulli_ys_numpy(B: int, T: int, seed: int, dt: float = 1e-2) -> np.ndarray:
    """uint8 (B, T) observations; NumPy ``Generator(PCG64(seed))``."""
    rng = np.random.Generator(np.random.PCG64(seed))
    comp = rng.integers(0, 2, B)
    x = np.where(comp == 0, -0.5, 0.5) + np.sqrt(0.05) * rng.standard_normal(B)
    ys = np.empty((B, T), dtype=np.uint8)
    for t in range(T):
        s = np.where(rng.random(B) < 0.5 * (1 + np.tanh(x)), 1., -1.)
        x = x + s * dt + np.sqrt(dt) * rng.standard_normal(B)
        with np.errstate(over='ignore'):
            ys[:, t] = rng.random(B) < 1 / (1 + np.exp(-x ** 3 / 5))
    return ys


def benes_bernoulli_ys_torch(B: int, T: int, seed: int, device, dt: float = 1e-2):
    """Same law generated on ``device`` with ``torch.Generator(seed)`` (for the 1e6 x 1000 bench batch).
    Returns a uint8 (B, T) tensor on ``device``."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device, generator=g)
    comp = torch.rand(B, **f64) < 0.5
    x = torch.where(comp, -0.5, 0.5) + (0.05 ** 0.5) * torch.randn(B, **f64)
    ys = torch.empty((B, T), dtype=torch.uint8, device=device)
    sq = dt ** 0.5
    for t in range(T):
        s = torch.where(torch.rand(B, **f64) < 0.5 * (1 + torch.tanh(x)), 1., -1.)
        x = x + s * dt + sq * torch.randn(B, **f64)
        ys[:, t] = (torch.rand(B, **f64) < torch.sigmoid(x ** 3 / 5)).to(torch.uint8)
    return ys


def ou_gaussian_ys_numpy(B: int, T: int, seed: int, dt: float = 0.1, ell: float = 1., sigma: float = 0.5,
                         r: float = 1.) -> np.ndarray:
    """float64 (B, T) observations of the OU + Gaussian model of ``dardel/convergence/convergence_mf.py:32-61``."""
    rng = np.random.Generator(np.random.PCG64(seed))
    F, Sig = np.exp(-dt / ell), sigma ** 2 * (1 - np.exp(-2 * dt / ell))
    x = sigma * rng.standard_normal(B)
    ys = np.empty((B, T))
    for t in range(T):
        x = F * x + np.sqrt(Sig) * rng.standard_normal(B)
        ys[:, t] = x + r * rng.standard_normal(B)
    return ys


def well_poisson_ys_numpy(B: int, T: int, seed: int, theta=(3., 3.), dt: float = 1e-2, substeps: int = 10) -> np.ndarray:
    """int32 (B, T) Poisson counts of the double-well model (``mfs/one_dim/ss_models.py:59-93``), Euler sub-stepping."""
    rng = np.random.Generator(np.random.PCG64(seed))
    comp = rng.integers(0, 2, B)
    x = np.where(comp == 0, -0.5, 0.5) + np.sqrt(0.05) * rng.standard_normal(B)
    ys = np.empty((B, T), dtype=np.int32)
    h = dt / substeps
    for t in range(T):
        for _ in range(substeps):
            x = x + x * (1 - theta[0] * x ** 2) * h + np.sqrt(h) * rng.standard_normal(B)
        ys[:, t] = rng.poisson(np.log1p(np.exp(theta[1] * x)))
    return ys

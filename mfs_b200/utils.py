"""Host-side mirror of the pieces of ``mfs/utils.py`` that sit on the filter path: Gaussian-sum initial moments."""
from typing import NamedTuple

import numpy as np

from .one_dim.moments import raw_moment_of_normal


class GaussianSum1D(NamedTuple):
    """Unidimensional Gaussian-sum distribution (``mfs/utils.py:39-74``); ``new`` computes rms / cms / scms."""
    means: np.ndarray
    variances: np.ndarray
    weights: np.ndarray
    mean: float
    variance: float
    rms: np.ndarray
    cms: np.ndarray
    scms: np.ndarray

    def pdf(self, xs):
        xs = np.atleast_1d(xs)[:, None]
        pdfs = np.exp(-0.5 * (xs - self.means) ** 2 / self.variances) / np.sqrt(2 * np.pi * self.variances)
        return np.sum(pdfs * self.weights[None, :], axis=1)

    def sampler(self, rng: np.random.Generator, n: int):
        cs = rng.choice(self.means.shape[0], (n,), p=self.weights)
        return self.means[cs] + np.sqrt(self.variances[cs]) * rng.standard_normal(n)

    @classmethod
    def new(cls, means, variances, weights, N: int = 2):
        means, variances, weights = (np.asarray(a, dtype=np.float64) for a in (means, variances, weights))
        centre = float(np.sum(means * weights))
        rms = np.array([sum(raw_moment_of_normal(m, v, p) * w for m, v, w in zip(means, variances, weights))
                        for p in range(2 * N)])
        cms = np.array([sum(raw_moment_of_normal(m - centre, v, p) * w for m, v, w in zip(means, variances, weights))
                        for p in range(2 * N)])
        variance = float(cms[2])
        scms = cms / np.sqrt(variance) ** np.arange(2 * N)
        return cls(means=means, variances=variances, weights=weights, mean=centre, variance=variance,
                   rms=rms, cms=cms, scms=scms)


class GaussianSumND(NamedTuple):
    """Multidimensional Gaussian-sum distribution (``mfs/utils.py:77-131``); ``new`` computes rms / cms."""
    d: int
    means: np.ndarray
    covs: np.ndarray
    weights: np.ndarray
    mean: np.ndarray
    cov: np.ndarray
    rms: np.ndarray
    cms: np.ndarray

    def sampler(self, rng: np.random.Generator, n: int):
        cs = rng.choice(self.means.shape[0], (n,), p=self.weights)
        chol = np.linalg.cholesky(self.covs)
        return self.means[cs] + np.einsum('nij,nj->ni', chol[cs], rng.standard_normal((n, self.d)))

    @classmethod
    def new(cls, means, covs, weights, multi_indices):
        from .multi_dims.moments import _gaussian_product_moments
        means, covs, weights = (np.asarray(a, dtype=np.float64) for a in (means, covs, weights))
        d = means.shape[1]
        centre = np.sum(means * weights[:, None], axis=0)
        cov = sum(w * (c + np.outer(m, m)) for m, c, w in zip(means, covs, weights)) - np.outer(centre, centre)
        rms = sum(w * _gaussian_product_moments(m, c, multi_indices) for m, c, w in zip(means, covs, weights))
        cms = sum(w * _gaussian_product_moments(m - centre, c, multi_indices) for m, c, w in zip(means, covs, weights))
        return cls(d=d, means=means, covs=covs, weights=weights, mean=centre, cov=cov, rms=rms, cms=cms)

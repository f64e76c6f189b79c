"""TEST INFRASTRUCTURE ONLY -- gradient oracle for the parameter-estimation objective (never imported by ``mfs_b200``).

The reference obtains d nell / d theta with ``jax.grad`` through the filter scan
(``dardel/parameter_estimation/mf.py:37-54``), which cannot run here (no JAX).  This module restates the reference's
DENSE algorithm for that objective --

  * ``moment_filter_cms``                    ``mfs/one_dim/filtering.py:129-161``
  * ``moment_quadrature``                    ``mfs/one_dim/quadtures.py:122-133`` (Hankel gather, Cholesky of G,
                                             K = R^-1 H R^-T, symmetric eigen-decomposition, w = V[0,:]^2)
  * ``sde_cond_moments_tme_normal`` (order 2) / ``sde_cond_moments_euler``   ``mfs/one_dim/moments.py:182-255``
  * ``well_poisson`` drift and pmf           ``mfs/one_dim/ss_models.py:70-84``

-- in arithmetic that is ANALYTIC in the parameters (plain complex numbers, no conjugation, no LAPACK: Cholesky by
the textbook recurrence, triangular solves by substitution, a cyclic Jacobi eigen-solver whose rotations are smooth
functions of the matrix entries and whose branches look at real parts only), so that the COMPLEX-STEP derivative
``Im nell(theta + i h) / h`` with h = 1e-30 is the exact derivative of the computed nell up to rounding: no
truncation error and no subtractive cancellation, unlike finite differences.  It is an independent route to the
gradient (dense algorithm, different eigen-solver, different differentiation technique) against which the CUDA
forward-mode kernel is compared; its VALUE is pinned to the NumPy / C oracles (and through them to the golden
vectors), its DERIVATIVE to central differences of those oracles (``tests/test_oracle_grad.py``).
Pure-Python loops: small cases only.
"""
import math

import numpy as np
from scipy.special import gammaln


def _cholesky(G):
    """Lower Cholesky factor by the textbook recurrence, analytic in the entries (no conjugation)."""
    n = G.shape[0]
    L = np.zeros((n, n), dtype=complex)
    for j in range(n):
        s = G[j, j] - np.sum(L[j, :j] ** 2)
        if not s.real > 0:
            return None
        L[j, j] = np.sqrt(s)
        for i in range(j + 1, n):
            L[i, j] = (G[i, j] - np.sum(L[i, :j] * L[j, :j])) / L[j, j]
    return L


def _solve_lower(L, B):
    """X with L X = B (forward substitution, column by column)."""
    n = L.shape[0]
    X = np.zeros_like(B, dtype=complex)
    for i in range(n):
        X[i] = (B[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def _jacobi_eigh(K, sweeps=14):
    """Cyclic Jacobi for a (complex-)symmetric matrix: A <- J^T A J with complex-orthogonal plane rotations
    (c^2 + s^2 = 1 holds identically), eigenvectors accumulated in V.  Branches use real parts only."""
    n = K.shape[0]
    A = K.astype(complex).copy()
    V = np.eye(n, dtype=complex)
    for _ in range(sweeps):
        for p in range(n - 1):
            for q in range(p + 1, n):
                apq = A[p, q]
                if abs(apq) < 1e-280:
                    continue
                theta = (A[q, q] - A[p, p]) / (2.0 * apq)
                if abs(theta) > 1e140:                       # converged entry: tan(angle) -> 1 / (2 theta)
                    t = 0.5 / theta
                else:
                    root = np.sqrt(theta * theta + 1.0)
                    t = 1.0 / (theta + (root if theta.real >= 0 else -root))
                c = 1.0 / np.sqrt(t * t + 1.0)
                s = t * c
                Ap, Aq = A[:, p].copy(), A[:, q].copy()
                A[:, p], A[:, q] = c * Ap - s * Aq, s * Ap + c * Aq
                Ap, Aq = A[p, :].copy(), A[q, :].copy()
                A[p, :], A[q, :] = c * Ap - s * Aq, s * Ap + c * Aq
                Vp, Vq = V[:, p].copy(), V[:, q].copy()
                V[:, p], V[:, q] = c * Vp - s * Vq, s * Vp + c * Vq
    return np.diag(A).copy(), V


def moment_quadrature(ms, mean):
    """quadtures.py:122-133 in analytic complex arithmetic -> (weights, nodes) or (None, None) when G is not PD."""
    n = len(ms) // 2
    idx = np.arange(n)[:, None] + np.arange(n)[None, :]
    G, H = ms[idx], ms[idx + 1]
    R = _cholesky(G)
    if R is None:
        return None, None
    Y = _solve_lower(R, H)                     # R^-1 H
    K = _solve_lower(R, Y.T).T                 # (R^-1 H) R^-T
    K = 0.5 * (K + K.T)
    lam, V = _jacobi_eigh(K)
    return V[0, :] ** 2, lam + mean


def _normal_moments(mu, var, num):
    """E[(mu + sqrt(var) xi)^p], p < num: the binomial sum of moments.py:70-74 without its vanishing odd terms."""
    out = []
    for p in range(num):
        acc = 0.0
        for m in range(p, -1, -2):             # p - m even
            k = p - m
            dfact = 1.0
            for j in range(k - 1, 0, -2):
                dfact *= j
            acc = acc + math.comb(p, m) * mu ** m * var ** (k // 2) * dfact
        out.append(acc)
    return out


def well_mean_var(x, theta1, dt, b=1.0, order=2):
    """tme.mean_and_cov for a(x) = x (1 - theta1 x^2), constant dispersion b (order 1 = Euler--Maruyama, order 2:
    mean = x + dt a + dt^2/2 (a a' + c a''), var = 2 c dt + dt^2/2 4 c a', c = b^2/2)."""
    c = 0.5 * b * b
    a0 = x * (1.0 - theta1 * x * x)
    a1 = 1.0 - 3.0 * theta1 * x * x
    a2 = -6.0 * theta1 * x
    mean = x + dt * a0
    var = 2.0 * c * dt + 0.0 * x
    if order >= 2:
        mean = mean + 0.5 * dt * dt * (a0 * a1 + c * a2)
        var = var + 0.5 * dt * dt * 4.0 * c * a1
    return mean, var


def poisson_softplus_pmf(y, x, theta2):
    mu = np.log(1.0 + np.exp(theta2 * x))
    if y == 0:
        return np.exp(-mu)
    return np.exp(y * np.log(mu) - gammaln(y + 1.0) - mu)


def well_poisson_nell(theta1, theta2, cms0, mean0, ys, dt=1e-2, order=2):
    """nell of moment_filter_cms for the well--Poisson model (central moments, Normal transition family); complex
    ``theta`` allowed.  Returns NaN when a Gram matrix loses positive definiteness (like the JAX scan)."""
    cms = np.asarray(cms0, dtype=complex).copy()
    mean = complex(mean0)
    num = len(cms)
    nell = 0.0 + 0.0j
    with np.errstate(all='ignore'):
        return _scan(theta1, theta2, cms, mean, num, nell, ys, dt, order)


def _scan(theta1, theta2, cms, mean, num, nell, ys, dt, order):
    for y in ys:
        w, x = moment_quadrature(cms, mean)                                   # filtering.py:145
        if w is None:
            return complex(np.nan)
        mu, var = well_mean_var(x, theta1, dt, order=order)
        mean = np.sum(w * mu)                                                  # :147
        mom = _normal_moments(mu - mean, var, num)
        cms = np.array([np.sum(w * mom[p]) for p in range(num)])              # :148
        w, x = moment_quadrature(cms, mean)                                   # :151
        if w is None:
            return complex(np.nan)
        lik = poisson_softplus_pmf(float(y), x, theta2)
        pdf_y = np.sum(w * lik)                                                # :152
        mean = np.sum(w * lik * x) / pdf_y                                     # :153
        cms = np.array([np.sum(w * lik * (x - mean) ** p) for p in range(num)]) / pdf_y   # :154-156
        nell = nell - np.log(pdf_y)
    return nell


def well_poisson_value_and_grad(theta1, theta2, cms0, mean0, ys, dt=1e-2, order=2, h=1e-30):
    """(nell, [d nell / d theta1, d nell / d theta2]) by the complex-step derivative."""
    f1 = well_poisson_nell(theta1 + 1j * h, theta2, cms0, mean0, ys, dt, order)
    f2 = well_poisson_nell(theta1, theta2 + 1j * h, cms0, mean0, ys, dt, order)
    return f1.real, np.array([f1.imag / h, f2.imag / h])

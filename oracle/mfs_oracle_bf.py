"""
CPU oracle for the brute-force grid filter  --  TEST INFRASTRUCTURE ONLY (see oracle/mfs_oracle.py for the rules).

NumPy restatement of ``mfs/classical_filters_smoothers/brute_force.py:26-136``.  The reference evaluates, per
integration sub-step, ``jnp.trapz(norm.pdf(x_i, m, scale) * ps, xs)`` for every grid point ``x_i`` (vmap, :121), i.e.
the trapezoidal rule applied row by row to the transition-density matrix ``P[i, j] = N(x_i; m_j, scale_j)``; then the
pointwise Bayes update normalised by another trapezoidal integral (:135).  This file keeps that operation order
(product first, trapezoid second) so that it is the thing the GEMM-shaped CUDA path is compared against.

Pinned by ``tests/golden/golden_brute_force.npz`` (the reference's own ``brute_force.py`` executed on the NumPy jax
shim: 'chapman-euler' and 'kolmogorov' are 100 % reference code; for 'chapman-tme-k' the shim's ``tme.mean_and_cov``
is the definition-driven restatement of ``oracle/mfs_oracle.py``) and by the reference's known-answer test
(``tests/test_classical_filters_smoothers.py:204-233``: OU + Gaussian likelihood against the Kalman filter).
"""
from __future__ import annotations

import numpy as np

from . import mfs_oracle as O


def trapz(y, x):
    """``jnp.trapz(y, x)`` along the last axis: 0.5 * sum(dx * (y[1:] + y[:-1]))."""
    dx = np.diff(x)
    return 0.5 * np.sum(dx * (y[..., 1:] + y[..., :-1]), axis=-1)


def gradient(f, dx):
    """``jnp.gradient(f, dx)``: second-order central differences inside, first-order one-sided at the two ends."""
    g = np.empty_like(f)
    g[1:-1] = (f[2:] - f[:-2]) / (2. * dx)
    g[0] = (f[1] - f[0]) / dx
    g[-1] = (f[-1] - f[-2]) / dx
    return g


def transition_mean_scale(drift_name, params, b, ddt, xs, pred_method):
    """brute_force.py:69-78: Euler--Maruyama, or tme.mean_and_cov of the given order, at every grid point."""
    xs = np.asarray(xs, dtype=np.float64)
    if pred_method == 'chapman-euler':
        a = drift_values(drift_name, params, xs)
        return xs + a * ddt, np.full_like(xs, b * np.sqrt(ddt))
    order = int(pred_method.split('-')[-1])
    _, mean_var = O.tme_1d(drift_name, params, b, ddt, order, 2)
    m, v = mean_var(xs)
    with np.errstate(invalid='ignore'):
        return m, np.sqrt(v)


def drift_values(drift_name, params, x):
    if drift_name == 'benes':
        return np.tanh(x)
    if drift_name == 'well':
        return x * (1. - params[0] * x ** 2)
    if drift_name == 'ou':
        return -x / params[0]
    if drift_name == 'linear':
        return params[0] * x
    raise ValueError(drift_name)


def drift_derivative(drift_name, params, x):
    if drift_name == 'benes':
        return 1. - np.tanh(x) ** 2
    if drift_name == 'well':
        return 1. - 3. * params[0] * x ** 2
    if drift_name == 'ou':
        return np.full_like(x, -1. / params[0])
    if drift_name == 'linear':
        return np.full_like(x, params[0])
    raise ValueError(drift_name)


def brute_force_filter(drift_name, params, b, measurement_cond_pdf, init_ps, xs, ys, dt, integration_steps=1,
                       pred_method='chapman-tme-2'):
    """brute_force.py:26-136 for a named drift with constant dispersion ``b``.  ``measurement_cond_pdf(y, xs)`` is a
    NumPy callable.  Returns (T, n)."""
    xs = np.asarray(xs, dtype=np.float64)
    ps = np.asarray(init_ps, dtype=np.float64).copy()
    dx = xs[1] - xs[0]
    ddt = dt / integration_steps
    out = np.empty((len(ys), xs.shape[0]))
    if 'chapman' in pred_method:
        m, scale = transition_mean_scale(drift_name, params, b, ddt, xs, pred_method)
        with np.errstate(all='ignore'):
            P = O.norm_pdf(xs[:, None], m[None, :], scale[None, :])     # P[i, j] = p(x_i | x_j)
    elif pred_method == 'kolmogorov':
        a, da = drift_values(drift_name, params, xs), drift_derivative(drift_name, params, xs)
        gam = b * b
    else:
        raise NotImplementedError(pred_method)
    for t, y in enumerate(ys):
        for _ in range(integration_steps):
            if 'chapman' in pred_method:
                ps = trapz(P * ps[None, :], xs)
            else:
                d1 = gradient(ps, dx)
                d2 = gradient(gradient(ps, dx), dx)
                ps = ps + (-(da * ps + a * d1) + 0.5 * gam * d2) * ddt   # dispersion constant: gamma' = gamma'' = 0
        lik = measurement_cond_pdf(y, xs)
        ps = lik * ps / trapz(lik * ps, xs)
        out[t] = ps
    return out

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's data simulators (never imported by ``mfs_b200``).

Follows, line by line,
  * ``simulate_sde``                                   ``mfs/utils.py:190-249`` (Gaussian sub-steps with ``m_and_cov``),
  * ``benes_bernoulli`` / ``well_poisson`` ``simulate_trajectory``   ``mfs/one_dim/ss_models.py:49-54, 86-91``
    (``tme.mean_and_cov(order=3)``, 100 sub-steps; TME from its definition, ``oracle.mfs_oracle.tme_1d``),
  * ``GaussianSum1D.sampler`` / ``GaussianSumND.sampler``            ``mfs/utils.py:55-58, 101-105``,
  * ``prey_predator(...).simulate`` (Milstein)                       ``mfs/multi_dims/ss_models.py:76-93``,
  * the measurement draws of ``dardel/benes_bernoulli/mf.py:80`` and ``dardel/parameter_estimation/mf.py:65``.

Random numbers: the reference uses ``jax.random`` keys (``rng_keys.npy``), not reproducible without JAX.  The product
defines its own counter-based stream -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), key = seed, counter =
(trajectory lo, trajectory hi, time index, draw index) -- restated here in NumPy and pinned to the published
known-answer vectors of the Random123 distribution (``tests/test_oracle_sim.py``).  Parity of the simulators is
therefore: bit-exact uniforms, states to rounding (libm vs CUDA ``log``/``sincospi``), measurements equal except where
a uniform falls within rounding of its threshold.  **The law is the reference's; the stream is not ("parity
unpinned" against jax.random).**
"""
import numpy as np

from . import mfs_oracle as O

DRAW_MEASUREMENT = 0x80000000
DRAW_SIGN = 0x80000001
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32-valued, key: (2,) -> (..., 4) uint32 (as uint64 arrays holding 32-bit values)."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) & _MASK for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1)


def draw(traj, time_index, j, seed):
    """Two uniforms (ua, ub) in (0, 1) per trajectory: u = (k + 1/2) 2^-52 with k the top 52 bits of a word pair."""
    traj = np.asarray(traj, dtype=np.uint64)
    ctr = np.stack([traj & _MASK, traj >> np.uint64(32), np.full(traj.shape, time_index, dtype=np.uint64),
                    np.full(traj.shape, j, dtype=np.uint64)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = lambda a, b: (((a << np.uint64(20)) | (b >> np.uint64(12))).astype(np.float64) + 0.5) * 2.0 ** -52
    return u(r[..., 0], r[..., 1]), u(r[..., 2], r[..., 3])


def normals(ua, ub):
    """Box--Muller."""
    r = np.sqrt(-2.0 * np.log(ua))
    return r * np.cos(2.0 * np.pi * ub), r * np.sin(2.0 * np.pi * ub)


def mixture_component(u, weights):
    cum = np.cumsum(np.asarray(weights, dtype=np.float64))[:-1]
    return np.sum(u[:, None] >= cum[None, :], axis=1)


def measure(meas, mp, x, ua, ub):
    if meas == 'bernoulli_logistic_cubic':
        with np.errstate(over='ignore'):
            p = 1.0 / (1.0 + np.exp(-(x ** 3 / mp[0] - mp[1])))
        return (ua < p).astype(np.float64)
    if meas == 'poisson_softplus':
        with np.errstate(over='ignore'):
            lam = np.log(1.0 + np.exp(mp[0] * x))
        big = ~(lam <= 64.0)                   # Normal approximation with continuity correction for large rates
        lam_s = np.where(big, 1.0, lam)
        p = np.exp(-lam_s)
        F = p.copy()
        k = np.zeros(x.shape)
        live = (ua > F) & ~big
        while live.any():
            k[live] += 1
            p[live] *= lam_s[live] / k[live]
            F[live] += p[live]
            live = live & (ua > F)
        if big.any():
            z0 = normals(ua, ub)[0]
            with np.errstate(invalid='ignore'):
                k = np.where(big, np.maximum(0.0, np.floor(lam + np.sqrt(lam) * z0 + 0.5)), k)
        return k
    if meas == 'gaussian':
        return mp[0] * x + mp[1] * normals(ua, ub)[0]
    raise ValueError(meas)


def simulate_1d(drift_name, drift_params, b, dt, T, means, variances, weights, meas, meas_params, B, seed,
                integration_steps=100, tme_order=3, scheme='tme', traj_offset=0):
    """-> x0 (B,), xs (B, T), ys (B, T) float64."""
    traj = np.arange(B, dtype=np.uint64) + np.uint64(traj_offset)
    k = mixture_component(draw(traj, 0, 0, seed)[0], weights)
    z0, _ = normals(*draw(traj, 0, 1, seed))
    x = np.asarray(means, dtype=np.float64)[k] + np.sqrt(np.asarray(variances, dtype=np.float64)[k]) * z0
    x0 = x.copy()
    xs, ys = np.empty((B, T)), np.empty((B, T))
    ddt = dt / integration_steps
    if scheme == 'tme':
        _, mean_var = O.tme_1d(drift_name, drift_params, b, ddt, tme_order, 2)
    for t in range(T):
        if scheme == 'benes_exact':
            z0, _ = normals(*draw(traj, t + 1, 0, seed))
            s = np.where(draw(traj, t + 1, DRAW_SIGN, seed)[0] < 0.5 * (1.0 + np.tanh(x)), 1.0, -1.0)
            x = x + s * dt + np.sqrt(dt) * z0
        else:
            for j in range(0, integration_steps, 2):
                z0, z1 = normals(*draw(traj, t + 1, j >> 1, seed))
                m, v = mean_var(x)
                with np.errstate(invalid='ignore'):
                    x = m + np.sqrt(v) * z0                     # 1x1 Cholesky (mfs/utils.py:237-240)
                if j + 1 < integration_steps:
                    m, v = mean_var(x)
                    with np.errstate(invalid='ignore'):
                        x = m + np.sqrt(v) * z1
        xs[:, t] = x
        ys[:, t] = measure(meas, meas_params, x, *draw(traj, t + 1, DRAW_MEASUREMENT, seed))
    return x0, xs, ys


def simulate_lv(params, dt, T, means, covs, weights, meas_params, B, seed, integration_steps=100, obs_dim=0,
                traj_offset=0):
    """Milstein Lotka--Volterra (mfs/multi_dims/ss_models.py:76-93) -> x0 (B, 2), xs (B, T, 2), ys (B, T)."""
    alp, beta, delta, gamma, sigma = params
    traj = np.arange(B, dtype=np.uint64) + np.uint64(traj_offset)
    k = mixture_component(draw(traj, 0, 0, seed)[0], weights)
    z = np.stack(normals(*draw(traj, 0, 1, seed)), axis=-1)
    L = np.linalg.cholesky(np.asarray(covs, dtype=np.float64))
    x = np.asarray(means, dtype=np.float64)[k] + np.einsum('...ij,...j->...i', L[k], z)
    x0 = x.copy()
    xs, ys = np.empty((B, T, 2)), np.empty((B, T))
    ddt = dt / integration_steps
    for t in range(T):
        for j in range(integration_steps):
            ddw = np.sqrt(ddt) * np.stack(normals(*draw(traj, t + 1, j, seed)), axis=-1)
            drift = x * (x[:, ::-1] * np.array([-beta, delta]) + np.array([alp, -gamma]))
            x = x + drift * ddt + sigma * x * ddw + 0.5 * sigma ** 2 * x * (ddw ** 2 - ddt)
        xs[:, t] = x
        ys[:, t] = measure('bernoulli_logistic_cubic', meas_params, x[:, obs_dim],
                           *draw(traj, t + 1, DRAW_MEASUREMENT, seed))
    return x0, xs, ys

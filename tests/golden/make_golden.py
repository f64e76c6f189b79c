"""
Generate the golden fixtures under ``tests/golden/`` by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (needs ``/root/reference``; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference is pure JAX and JAX is not installed, so its modules are imported on top of the NumPy-backed shim in
``_jax_shim.py`` (float64, LAPACK potrf/trtrs/syevd = the routines XLA:CPU dispatches to).  What runs is the
reference's code, unmodified, from where it lies:

  mfs/one_dim/quadtures.py   moment_quadrature            -> golden_quadrature_1d.npz
  mfs/one_dim/filtering.py   moment_filter_{rms,cms,scms} -> golden_filter_1d_*.npz
  mfs/one_dim/moments.py     raw_to_central, raw_to_scaled, sde_cond_moments_euler, raw_moment_of_normal
  mfs/one_dim/ss_models.py   benes_bernoulli (constants, logistic, Bernoulli pmf), well_poisson
  mfs/utils.py               GaussianSum1D.new, ldl, ldl_chol
  mfs/multi_dims/multi_indices.py   (pure NumPy)          -> golden_multi_indices.npz
  mfs/multi_dims/moments.py  Kan--Magnus NumPy branches   -> golden_kan_moments.npz
  mfs/multi_dims/quadratures.py, filtering.py, ss_models.py (prey_predator), utils.GaussianSumND -> golden_nd.npz
  mfs/classical_filters_smoothers/brute_force.py  brute_force_filter -> golden_brute_force.npz
  mfs/one_dim/moments.py characteristic_fn -> golden_characteristic.npz
  mfs/utils.py ldl_chol through moment_quadrature(ldl=True) / moment_filter_rms(stable=True) -> golden_stable.npz

The third-party ``tme`` package is absent, so fixtures whose transition moments are TME expansions take those
callables from ``oracle/mfs_oracle.py`` (definition-driven restatement) and are labelled ``*_tme*``: they pin the
filter recursion + quadrature of the reference, not ``tme`` itself.  Fixtures labelled ``*_euler*`` are 100 %
reference code (Euler--Maruyama + Normal transition, ``mfs/one_dim/moments.py:222-255``).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('MFS_REFERENCE', '/root/reference')
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _jax_shim  # noqa: E402

_jax_shim.install()
sys.path.insert(0, REF)

import jax.numpy as jnp  # noqa: E402  (the shim)
from mfs.one_dim.quadtures import moment_quadrature  # noqa: E402
from mfs.one_dim.filtering import moment_filter_rms, moment_filter_cms, moment_filter_scms  # noqa: E402
from mfs.one_dim.moments import (raw_to_central, raw_to_scaled, raw_moment_of_normal,  # noqa: E402
                                 sde_cond_moments_euler, central_to_raw)
from mfs.one_dim.ss_models import benes_bernoulli, well_poisson  # noqa: E402
from mfs.utils import GaussianSum1D, ldl_chol  # noqa: E402

from oracle import mfs_oracle as O  # noqa: E402


def synth_benes_bernoulli(rng, T, dt, n_traj):
    """Exact Benes law (mixture of two Normals) + Bernoulli(logistic(x^3/5)) observations, uint8 (n_traj, T)."""
    comp = rng.integers(0, 2, n_traj)
    x = np.where(comp == 0, -0.5, 0.5) + np.sqrt(0.05) * rng.standard_normal(n_traj)
    ys = np.empty((n_traj, T), dtype=np.uint8)
    for t in range(T):
        s = np.where(rng.random(n_traj) < 0.5 * (1 + np.tanh(x)), 1., -1.)
        x = x + s * dt + np.sqrt(dt) * rng.standard_normal(n_traj)
        ys[:, t] = rng.random(n_traj) < 1 / (1 + np.exp(-x ** 3 / 5))
    return ys


def golden_quadrature_1d():
    out = {}
    cases = {}
    # Gaussian m=0.2, v=1.1, n=8 nodes  (tests/test_one_dim_quadrature.py:50-57 setting)
    rms = np.array([raw_moment_of_normal(0.2, 1.1, p) for p in range(16)])
    cases['gauss8_raw'] = (rms, 0., 1.)
    cases['gauss8_central'] = (np.asarray(raw_to_central(jnp.asarray(rms))), 0.2, 1.)
    cases['gauss8_scaled'] = (np.asarray(raw_to_scaled(jnp.asarray(rms))), 0.2, np.sqrt(1.1))
    for N in (2, 3, 5, 8, 10, 12, 15):
        ic = GaussianSum1D.new(means=jnp.array([-0.5, 0.5]), variances=jnp.array([0.05, 0.05]),
                               weights=jnp.array([0.5, 0.5]), N=N)
        cases[f'mix{N}_raw'] = (np.asarray(ic.rms), 0., 1.)
        cases[f'mix{N}_central'] = (np.asarray(ic.cms), float(ic.mean), 1.)
        cases[f'mix{N}_scaled'] = (np.asarray(ic.scms), float(ic.mean), float(np.sqrt(ic.variance)))
    # uniform[-2, 3] moments, N=4 (tests/test_one_dim_quadrature.py:96-113 setting)
    a, b = -2., 3.
    cases['uniform4_raw'] = (np.array([(b ** (p + 1) - a ** (p + 1)) / ((p + 1) * (b - a)) for p in range(8)]), 0., 1.)
    for name, (ms, mean, scale) in cases.items():
        w, x = moment_quadrature(jnp.asarray(ms), mean, scale, sort_nodes=True)
        out[f'{name}/ms'], out[f'{name}/mean'], out[f'{name}/scale'] = ms, mean, scale
        out[f'{name}/weights'], out[f'{name}/nodes'] = np.asarray(w), np.asarray(x)
        w, x = moment_quadrature(jnp.asarray(ms), mean, scale, sort_nodes=True, ldl=True)
        out[f'{name}/weights_ldl'], out[f'{name}/nodes_ldl'] = np.asarray(w), np.asarray(x)
    # A non-PD Hankel matrix: NaN semantics of the reference, and the ldl completion
    bad = np.array([1., 0., 1., 0., 0.5, 0., 3., 0.])
    w, x = moment_quadrature(jnp.asarray(bad))
    out['nonpd/ms'], out['nonpd/weights'], out['nonpd/nodes'] = bad, np.asarray(w), np.asarray(x)
    G = bad[np.arange(4)[:, None] + np.arange(4)[None, :]]
    out['nonpd/ldl_chol'] = np.asarray(ldl_chol(jnp.asarray(G)))
    np.savez_compressed(os.path.join(HERE, 'golden_quadrature_1d.npz'), **out)
    print('golden_quadrature_1d.npz', len(cases), 'cases')


def golden_conversions():
    out = {}
    for N in (2, 5, 8):
        ic = GaussianSum1D.new(means=jnp.array([-0.5, 0.5]), variances=jnp.array([0.05, 0.05]),
                               weights=jnp.array([0.5, 0.5]), N=N)
        out[f'mix{N}/rms'], out[f'mix{N}/cms'], out[f'mix{N}/scms'] = map(np.asarray, (ic.rms, ic.cms, ic.scms))
        out[f'mix{N}/mean'], out[f'mix{N}/variance'] = float(ic.mean), float(ic.variance)
    rms = np.array([raw_moment_of_normal(1.1, 4.9, p) for p in range(10)])
    out['normal/rms'] = rms
    out['normal/raw_to_central'] = np.asarray(raw_to_central(jnp.asarray(rms)))
    out['normal/raw_to_scaled'] = np.asarray(raw_to_scaled(jnp.asarray(rms)))
    out['normal/central_to_raw'] = np.asarray(central_to_raw(jnp.asarray(out['normal/raw_to_central']), 1.1))
    np.savez_compressed(os.path.join(HERE, 'golden_conversions_1d.npz'), **out)
    print('golden_conversions_1d.npz')


def golden_filter_1d():
    rng = np.random.Generator(np.random.PCG64(666))
    n_traj = 4
    for N in (5, 8):
        dt, T, ts, init_cond, drift, dispersion, logistic, pmf, _ = benes_bernoulli(N)
        ys_all = synth_benes_bernoulli(rng, T, dt, n_traj)
        out = {'ys': ys_all, 'rms0': np.asarray(init_cond.rms), 'cms0': np.asarray(init_cond.cms),
               'scms0': np.asarray(init_cond.scms), 'mean0': float(init_cond.mean),
               'scale0': float(np.sqrt(init_cond.variance)), 'dt': dt}

        # (a) 100 % reference code: Euler--Maruyama + Normal transition moments
        e_rms, e_cms, e_scms, e_mean, e_mean_var = sde_cond_moments_euler(drift, dispersion, dt, N)
        # (b) reference filter + quadrature, TME-2 / TME-3 transition moments from the oracle restatement
        variants = {'euler': (e_rms, e_cms, e_scms, e_mean, e_mean_var)}
        for order in (2, 3):
            variants[f'tme{order}'] = O.sde_cond_moments_tme('benes', (), 1., dt, order, 2 * N)
        variants['tme_normal3'] = O.sde_cond_moments_tme_normal('benes', (), 1., dt, 3, N)

        for vname, (f_rms, f_cms, f_scms, f_mean, f_mean_var) in variants.items():
            for k in range(n_traj):
                ys = jnp.asarray(ys_all[k])
                rmss, nell = moment_filter_rms(f_rms, pmf, init_cond.rms, ys)
                out[f'{vname}/rms/{k}/rmss'], out[f'{vname}/rms/{k}/nell'] = np.asarray(rmss), float(nell)
                cmss, means, nell = moment_filter_cms(f_cms, f_mean, pmf, init_cond.cms, init_cond.mean, ys)
                out[f'{vname}/cms/{k}/cmss'], out[f'{vname}/cms/{k}/means'] = np.asarray(cmss), np.asarray(means)
                out[f'{vname}/cms/{k}/nell'] = float(nell)
                if vname in ('tme2', 'tme3'):   # the Normal factories' scaled variant divides by prod(scale**k): skip
                    scmss, means, scales, nell = moment_filter_scms(f_scms, f_mean_var, pmf, init_cond.scms,
                                                                    init_cond.mean, jnp.sqrt(init_cond.variance), ys)
                    out[f'{vname}/scms/{k}/scmss'] = np.asarray(scmss)
                    out[f'{vname}/scms/{k}/means'], out[f'{vname}/scms/{k}/scales'] = map(np.asarray, (means, scales))
                    out[f'{vname}/scms/{k}/nell'] = float(nell)
            # stable=True (LDL completion) on the first trajectory
            rmss, nell = moment_filter_rms(f_rms, pmf, init_cond.rms, jnp.asarray(ys_all[0]), stable=True)
            out[f'{vname}/rms_stable/0/rmss'], out[f'{vname}/rms_stable/0/nell'] = np.asarray(rmss), float(nell)
        np.savez_compressed(os.path.join(HERE, f'golden_filter_1d_benes_N{N}.npz'), **out)
        print(f'golden_filter_1d_benes_N{N}.npz')

    # OU + Gaussian likelihood, the reference's own test input (tests/test_filtering.py:18-37; numpy seed 666)
    import math
    np.random.seed(666)
    dt, T = 1e-2, 100
    ts = np.linspace(dt, dt * T, T)
    ell, sigma = 1., 0.5
    cov = np.exp(-np.abs(ts[None, :] - ts[:, None]) / ell) * sigma ** 2
    ys = np.linalg.cholesky(cov) @ np.random.randn(T) + np.random.randn(T)
    b = math.sqrt(2) * sigma / math.sqrt(ell)
    from jax.scipy.stats import norm
    pdf = lambda y, x: jnp.squeeze(norm.pdf(y, x, 1.))
    out = {'ys': ys, 'ell': ell, 'sigma': sigma, 'dt': dt}
    for N, order, mean0, var0 in ((10, 3, 0.1, 0.1), (4, 2, 0., 0.5)):
        rms0 = jnp.array([raw_moment_of_normal(mean0, var0, p) for p in range(2 * N)])
        f_rms, f_cms, f_scms, f_mean, f_mean_var = O.sde_cond_moments_tme('ou', (ell,), b, dt, order, 2 * N)
        rmss, nell = moment_filter_rms(f_rms, pdf, rms0, jnp.asarray(ys))
        out[f'N{N}/rms0'], out[f'N{N}/rmss'], out[f'N{N}/nell'] = np.asarray(rms0), np.asarray(rmss), float(nell)
        cmss, means, nell = moment_filter_cms(f_cms, f_mean, pdf, raw_to_central(rms0), mean0, jnp.asarray(ys))
        out[f'N{N}/cmss'], out[f'N{N}/means'], out[f'N{N}/nell_c'] = np.asarray(cmss), np.asarray(means), float(nell)
        out[f'N{N}/mean0'], out[f'N{N}/var0'], out[f'N{N}/order'] = mean0, var0, order
    np.savez_compressed(os.path.join(HERE, 'golden_filter_1d_ou.npz'), **out)
    print('golden_filter_1d_ou.npz')

    # well--Poisson, central, Euler + the oracle's TME-normal-2  (dardel/parameter_estimation/mf.py:50-53 setting)
    N = 5
    dt, T, ts, init_cond, drift, dispersion, emission, pmf, _ = well_poisson(3., N)
    T = 200
    theta = (3., 3.)
    rng = np.random.Generator(np.random.PCG64(667))
    x = 0.5
    ys = np.empty(T, dtype=np.int32)
    for t in range(T):
        for _ in range(10):
            x = x + x * (1 - theta[0] * x ** 2) * dt / 10 + np.sqrt(dt / 10) * rng.standard_normal()
        ys[t] = rng.poisson(np.log1p(np.exp(theta[1] * x)))
    out = {'ys': ys, 'theta': np.array(theta), 'cms0': np.asarray(init_cond.cms), 'mean0': float(init_cond.mean),
           'rms0': np.asarray(init_cond.rms), 'dt': dt}
    e = sde_cond_moments_euler(lambda u: drift(u, theta[0]), dispersion, dt, N)
    tn = O.sde_cond_moments_tme_normal('well', (theta[0],), 1., dt, 2, N)
    for vname, fam in (('euler', e), ('tme_normal2', tn)):
        cmss, means, nell = moment_filter_cms(fam[1], fam[3], lambda y, u: pmf(y, u, theta[1]), init_cond.cms,
                                              init_cond.mean, jnp.asarray(ys))
        out[f'{vname}/cmss'], out[f'{vname}/means'], out[f'{vname}/nell'] = np.asarray(cmss), np.asarray(means), \
            float(nell)
        rmss, nell = moment_filter_rms(fam[0], lambda y, u: pmf(y, u, theta[1]), init_cond.rms, jnp.asarray(ys))
        out[f'{vname}/rmss'], out[f'{vname}/nell_r'] = np.asarray(rmss), float(nell)
    np.savez_compressed(os.path.join(HERE, 'golden_filter_1d_well.npz'), **out)
    print('golden_filter_1d_well.npz')


def synth_prey_predator(rng, T, dt, substeps=100):
    """Milstein simulation of the Lotka--Volterra SDE + Bernoulli observations (mfs/multi_dims/ss_models.py:69-93),
    NumPy RNG."""
    alp, beta, delta, gamma, sigma = 4., 4., 4., 4., 0.1
    x = np.array([1., 1.]) + np.sqrt(0.0015) * rng.standard_normal(2)
    ddt = dt / substeps
    xs = np.empty((T, 2))
    for t in range(T):
        for _ in range(substeps):
            ddw = np.sqrt(ddt) * rng.standard_normal(2)
            drift = x * (x[::-1] * np.array([-beta, delta]) + np.array([alp, -gamma]))
            x = x + drift * ddt + sigma * x * ddw + 0.5 * sigma ** 2 * x * (ddw ** 2 - ddt)
        xs[t] = x
    ys = (rng.random(T) < 1 / (1 + np.exp(-xs[:, 0] ** 3 + 1))).astype(np.uint8)
    return xs, ys


def golden_multi_indices():
    from mfs.multi_dims.multi_indices import (generate_graded_lexico_multi_indices,
                                              gram_and_hankel_indices_graded_lexico)
    out = {}
    for d in (1, 2, 3, 5):
        for up in (4, 6):
            for lo in (0, 3):
                out[f'gen/{d}_{up}_{lo}'] = generate_graded_lexico_multi_indices(d, up, lo)
    for N, d in ((2, 2), (3, 2), (5, 2), (8, 2), (3, 3), (4, 1), (2, 4)):
        out[f'gh/{N}_{d}'] = gram_and_hankel_indices_graded_lexico(N, d)
    np.savez_compressed(os.path.join(HERE, 'golden_multi_indices.npz'), **out)
    print('golden_multi_indices.npz')


def golden_nd():
    from mfs.multi_dims.multi_indices import (generate_graded_lexico_multi_indices,
                                              gram_and_hankel_indices_graded_lexico)
    from mfs.multi_dims.quadratures import moment_quadrature_nd
    from mfs.multi_dims.moments import raw_moments_mvn_kan, sde_cond_moments_euler_maruyama
    from mfs.multi_dims.filtering import moment_filter_nd_rms, moment_filter_nd_cms
    from mfs.multi_dims.ss_models import prey_predator
    d = 2
    out = {}
    # Kan--Magnus moments + nd quadrature of a correlated Gaussian (tests/test_multi_dim_quadrature.py setting)
    mean = np.array([0.3, -0.2])
    cov = np.array([[1.1, 0.3], [0.3, 0.7]])
    for N in (3, 4, 5):
        mis = generate_graded_lexico_multi_indices(d, 2 * N - 1, 0)
        inds = gram_and_hankel_indices_graded_lexico(N, d)
        rms = np.array([raw_moments_mvn_kan(mean, cov, n) for n in mis])
        cms = np.array([raw_moments_mvn_kan(mean - mean, cov, n) for n in mis])
        w, x = moment_quadrature_nd(jnp.asarray(rms), inds)
        wc, xc = moment_quadrature_nd(jnp.asarray(cms), inds, jnp.asarray(mean))
        out[f'quad/N{N}/rms'], out[f'quad/N{N}/cms'] = rms, cms
        out[f'quad/N{N}/w'], out[f'quad/N{N}/x'] = np.asarray(w), np.asarray(x)
        out[f'quad/N{N}/wc'], out[f'quad/N{N}/xc'] = np.asarray(wc), np.asarray(xc)
    out['quad/mean'], out['quad/cov'] = mean, cov
    # prey--predator filters, Euler--Maruyama + Normal transition (100 % reference code, Kan--Magnus NumPy branch)
    rng = np.random.Generator(np.random.PCG64(671))
    for N, T in ((3, 12), (4, 6)):
        mis = generate_graded_lexico_multi_indices(d, 2 * N - 1, 0)
        inds = gram_and_hankel_indices_graded_lexico(N, d)
        dt, _, ts, gs, drift, dispersion, emission, pmf, _ = prey_predator(mis)
        xs, ys = synth_prey_predator(rng, T, dt)
        f_rms, f_cms, _, f_mean, _ = sde_cond_moments_euler_maruyama(drift, dispersion, dt, mis)
        rmss, nell = moment_filter_nd_rms((f_rms, 'index'), pmf, jnp.asarray(ys), (jnp.asarray(mis), inds), gs.rms)
        cmss, means, nell_c = moment_filter_nd_cms((f_cms, 'index'), f_mean, pmf, jnp.asarray(ys),
                                                   (jnp.asarray(mis), inds), gs.cms, gs.mean)
        out[f'pp/N{N}/ys'], out[f'pp/N{N}/rms0'], out[f'pp/N{N}/cms0'] = ys, np.asarray(gs.rms), np.asarray(gs.cms)
        out[f'pp/N{N}/mean0'] = np.asarray(gs.mean)
        out[f'pp/N{N}/rmss'], out[f'pp/N{N}/nell'] = np.asarray(rmss), float(nell)
        out[f'pp/N{N}/cmss'], out[f'pp/N{N}/means'], out[f'pp/N{N}/nell_c'] = np.asarray(cmss), np.asarray(means), \
            float(nell_c)
    # initial moments of the paper setting N = 5
    mis = generate_graded_lexico_multi_indices(d, 9, 0)
    dt, _, ts, gs, *_ = prey_predator(mis)
    out['pp/N5/rms0'], out['pp/N5/cms0'], out['pp/N5/mean0'] = np.asarray(gs.rms), np.asarray(gs.cms), np.asarray(gs.mean)
    np.savez_compressed(os.path.join(HERE, 'golden_nd.npz'), **out)
    print('golden_nd.npz')


def golden_brute_force():
    """mfs/classical_filters_smoothers/brute_force.py executed on the shim.  'chapman-euler' and 'kolmogorov' are 100 %
    reference code (jax.grad of the shim = complex-step derivative, exact to rounding for analytic drifts); for
    'chapman-tme-k' the absent third-party ``tme.mean_and_cov`` is the oracle's definition-driven restatement."""
    import jax
    import tme.base_jax as tme_base
    from mfs.classical_filters_smoothers.brute_force import brute_force_filter
    from jax.scipy.stats import norm

    def grad(f):
        def df(x):
            h = 1e-30
            return np.imag(np.asarray(f(np.asarray(x, dtype=np.complex128) + 1j * h))) / h
        return df

    jax.grad = grad
    state = {}

    def mean_and_cov(x, ddt, drift, dispersion, order):
        _, mv = O.tme_1d(state['drift'], state['params'], state['b'], float(ddt), order, 2)
        m, v = mv(np.asarray(x))
        return m, v.reshape(1, 1)

    tme_base.mean_and_cov = mean_and_cov
    out = {}
    rng = np.random.Generator(np.random.PCG64(672))
    # (1) Benes--Bernoulli on a 200-point grid
    dt, T, ts, init_cond, drift, dispersion, logistic, pmf, _ = benes_bernoulli(3)
    state.update(drift='benes', params=(), b=1.)
    xs = np.linspace(-4., 4., 200)
    ys = synth_benes_bernoulli(rng, 6, dt, 3)
    init_ps = np.asarray(init_cond.pdf(jnp.asarray(xs)))
    out['benes/xs'], out['benes/ys'], out['benes/init_ps'], out['benes/dt'] = xs, ys, init_ps, dt
    for method, steps in (('chapman-euler', 5), ('chapman-tme-2', 5), ('chapman-tme-3', 5), ('chapman-tme-3', 1),
                          ('kolmogorov', 20)):
        for k in range(ys.shape[0]):
            pss = brute_force_filter(drift, dispersion, pmf, jnp.asarray(init_ps), jnp.asarray(xs), jnp.asarray(ys[k]),
                                     dt, integration_steps=steps, pred_method=method)
            out[f'benes/{method}/{steps}/{k}'] = np.asarray(pss)
    # (2) OU + Gaussian likelihood, the setting of tests/test_classical_filters_smoothers.py:127-160 on a smaller grid
    import math
    ell, sigma, r2 = 1., 0.5, 0.1
    b = math.sqrt(2) * sigma / math.sqrt(ell)
    state.update(drift='ou', params=(ell,), b=b)
    xs = np.linspace(-5., 5., 300)
    T = 10
    traj = np.cumsum(0.1 * rng.standard_normal(T))
    ys = traj + math.sqrt(r2) * rng.standard_normal(T)
    init_ps = np.asarray(norm.pdf(xs, 0., sigma))
    out['ou/xs'], out['ou/ys'], out['ou/init_ps'], out['ou/dt'] = xs, ys, init_ps, 1e-2
    out['ou/ell'], out['ou/sigma'], out['ou/r2'] = ell, sigma, r2
    for method, steps in (('chapman-euler', 4), ('chapman-tme-3', 4), ('kolmogorov', 20)):
        pss = brute_force_filter(lambda u: -1 / ell * u, lambda _: b, lambda y, x: norm.pdf(y, x, math.sqrt(r2)),
                                 jnp.asarray(init_ps), jnp.asarray(xs), jnp.asarray(ys), 1e-2,
                                 integration_steps=steps, pred_method=method)
        out[f'ou/{method}/{steps}'] = np.asarray(pss)
    np.savez_compressed(os.path.join(HERE, 'golden_brute_force.npz'), **out)
    print('golden_brute_force.npz')


def golden_stable():
    """`stable=True` / `ldl=True` on moment sequences whose Hankel matrix has NEGATIVE LDL pivots (the case where
    mfs/utils.py:538 substitutes eps and K stops being tridiagonal), 100 % reference code:
    quadratures (quadtures.py:127 with ldl=True) and an Euler--Maruyama raw-moment filter started from a non-PD
    initial sequence (filtering.py:73-86 with stable=True)."""
    out = {}
    cases = {'n4': np.array([1., 0., 1., 0., 0.5, 0., 3., 0.])}
    for N in (3, 5, 8):
        ic = GaussianSum1D.new(means=jnp.array([-0.5, 0.5]), variances=jnp.array([0.05, 0.05]),
                               weights=jnp.array([0.5, 0.5]), N=N)
        bad = np.asarray(ic.rms).copy()
        bad[2 * (N // 2)] *= 0.5          # breaks a middle pivot
        cases[f'mix{N}'] = bad
        bad2 = np.asarray(ic.rms).copy()
        bad2[2 * N - 2] *= 0.2            # breaks the last pivot only
        cases[f'mix{N}_last'] = bad2
    for name, ms in cases.items():
        w, x = moment_quadrature(jnp.asarray(ms), 0.3, 1.7, sort_nodes=True, ldl=True)
        out[f'quad/{name}/ms'], out[f'quad/{name}/weights'], out[f'quad/{name}/nodes'] = ms, np.asarray(w), np.asarray(x)
    rng = np.random.Generator(np.random.PCG64(673))
    N = 5
    dt, T, ts, init_cond, drift, dispersion, logistic, pmf, _ = benes_bernoulli(N)
    ys_all = synth_benes_bernoulli(rng, 30, dt, 2)
    e_rms, e_cms, e_scms, e_mean, e_mean_var = sde_cond_moments_euler(drift, dispersion, dt, N)
    bad = np.asarray(init_cond.rms).copy()
    bad[4] *= 0.9          # one negative LDL pivot at step 0; the run stays finite
    out['filter/ys'], out['filter/rms0'], out['filter/dt'] = ys_all, bad, dt
    for k in range(2):
        rmss, nell = moment_filter_rms(e_rms, pmf, jnp.asarray(bad), jnp.asarray(ys_all[k]), stable=True)
        out[f'filter/{k}/rmss'], out[f'filter/{k}/nell'] = np.asarray(rmss), float(nell)
    # d = 2: moment_quadrature_nd(ldl=True) on a Gram matrix with a negative pivot (quadratures.py:152)
    from mfs.multi_dims.multi_indices import (generate_graded_lexico_multi_indices,
                                              gram_and_hankel_indices_graded_lexico)
    from mfs.multi_dims.quadratures import moment_quadrature_nd
    from mfs.multi_dims.moments import raw_moments_mvn_kan
    mean, cov = np.array([0.3, -0.2]), np.array([[1.1, 0.3], [0.3, 0.7]])
    for Nn in (2, 3):
        mis = generate_graded_lexico_multi_indices(2, 2 * Nn - 1, 0)
        inds = gram_and_hankel_indices_graded_lexico(Nn, 2)
        rms = np.array([raw_moments_mvn_kan(mean, cov, n) for n in mis])
        bad = rms.copy()
        bad[int(np.where((mis == [2, 0]).all(axis=1))[0][0])] *= 0.2          # E[x1^2] too small: negative pivot
        w, x = moment_quadrature_nd(jnp.asarray(bad), inds, ldl=True)
        out[f'nd/N{Nn}/ms'], out[f'nd/N{Nn}/w'], out[f'nd/N{Nn}/x'] = bad, np.asarray(w), np.asarray(x)
        w, x = moment_quadrature_nd(jnp.asarray(rms), inds, ldl=True)
        out[f'nd/N{Nn}/ms_pd'], out[f'nd/N{Nn}/w_pd'], out[f'nd/N{Nn}/x_pd'] = rms, np.asarray(w), np.asarray(x)
    np.savez_compressed(os.path.join(HERE, 'golden_stable.npz'), **out)
    print('golden_stable.npz', {k: out[k] for k in out if k.endswith('nell')})


def golden_characteristic():
    """mfs/one_dim/moments.py:309-337 characteristic_fn on the shim (100 % reference code): raw / central / scaled
    moments of the Gaussian mixture, and on filter output (golden_filter_1d_benes_N5 raw moments)."""
    from mfs.one_dim.moments import characteristic_fn
    out = {}
    zs = np.linspace(-6., 6., 41)
    out['zs'] = zs
    for N in (3, 5, 8):
        ic = GaussianSum1D.new(means=jnp.array([-0.5, 0.5]), variances=jnp.array([0.05, 0.05]),
                               weights=jnp.array([0.5, 0.5]), N=N)
        for name, (ms, mean, scale) in {'raw': (ic.rms, 0., 1.), 'central': (ic.cms, float(ic.mean), 1.),
                                        'scaled': (ic.scms, float(ic.mean), float(np.sqrt(ic.variance)))}.items():
            out[f'mix{N}/{name}/ms'], out[f'mix{N}/{name}/mean'], out[f'mix{N}/{name}/scale'] = np.asarray(ms), mean, scale
            out[f'mix{N}/{name}/cf'] = np.array([complex(characteristic_fn(z, jnp.asarray(ms), mean, scale)) for z in zs])
    g = np.load(os.path.join(HERE, 'golden_filter_1d_benes_N5.npz'))
    rmss = g['tme3/rms/0/rmss'][::10]
    out['filter/rmss'] = rmss
    out['filter/cf'] = np.array([[complex(characteristic_fn(z, jnp.asarray(r))) for z in zs] for r in rmss])
    np.savez_compressed(os.path.join(HERE, 'golden_characteristic.npz'), **out)
    print('golden_characteristic.npz')


def golden_scaled_normal():
    """moment_filter_scms with the reference's Normal-approximation factories, whose scaled transition moments divide
    EVERY order by prod(scale ** arange(2N)) (mfs/one_dim/moments.py:205, 243).  The 'euler' entries are 100 % reference
    code; 'tme_normal3' takes tme.mean_and_cov from the oracle and the reference's own factory formula around it."""
    rng = np.random.Generator(np.random.PCG64(671))
    n_traj = 4
    out = {}
    for N in (3, 5):
        dt, T, ts, init_cond, drift, dispersion, logistic, pmf, _ = benes_bernoulli(N)
        ys_all = synth_benes_bernoulli(rng, T, dt, n_traj)
        out[f'N{N}/ys'], out[f'N{N}/scms0'] = ys_all, np.asarray(init_cond.scms)
        out[f'N{N}/mean0'], out[f'N{N}/scale0'], out[f'N{N}/dt'] = float(init_cond.mean), float(np.sqrt(init_cond.variance)), dt
        variants = {'euler': sde_cond_moments_euler(drift, dispersion, dt, N),
                    'tme_normal3': O.sde_cond_moments_tme_normal('benes', (), 1., dt, 3, N)}
        for vname, fam in variants.items():
            for k in range(n_traj):
                scmss, means, scales, nell = moment_filter_scms(fam[2], fam[4], pmf, init_cond.scms, init_cond.mean,
                                                                jnp.sqrt(init_cond.variance), jnp.asarray(ys_all[k]))
                out[f'N{N}/{vname}/{k}/scmss'], out[f'N{N}/{vname}/{k}/means'] = np.asarray(scmss), np.asarray(means)
                out[f'N{N}/{vname}/{k}/scales'], out[f'N{N}/{vname}/{k}/nell'] = np.asarray(scales), float(nell)
    np.savez_compressed(os.path.join(HERE, 'golden_filter_1d_scaled_normal.npz'), **out)
    print('golden_filter_1d_scaled_normal.npz')


if __name__ == '__main__':
    which = sys.argv[1:] or ['quadrature', 'conversions', 'filter1d', 'multi_indices', 'nd', 'brute_force', 'stable', 'characteristic', 'scaled_normal']
    if 'multi_indices' in which:
        golden_multi_indices()
    if 'nd' in which:
        golden_nd()
    if 'brute_force' in which:
        golden_brute_force()
    if 'stable' in which:
        golden_stable()
    if 'characteristic' in which:
        golden_characteristic()
    if 'quadrature' in which:
        golden_quadrature_1d()
    if 'conversions' in which:
        golden_conversions()
    if 'filter1d' in which:
        golden_filter_1d()
    if 'scaled_normal' in which:
        golden_scaled_normal()

"""Device data simulators: the step before the filter (``dardel/benes_bernoulli/mf.py:73-80``,
``dardel/parameter_estimation/mf.py:58-65``, ``mfs/multi_dims/ss_models.py:69-93``).

``simulate_1d`` runs ``init_cond.sampler`` + ``simulate_sde(tme.mean_and_cov(order), ..., integration_steps)`` +
the Bernoulli / Poisson / Gaussian measurement draw for a whole batch of trajectories in one kernel
(``mfs_simulate_1d``), ``simulate_prey_predator`` the Milstein Lotka--Volterra simulator (``mfs_simulate_lv``).
Random numbers come from the library's counter-based Philox4x32-10 stream (key = ``seed``, counter = global trajectory
id, time step, draw): ``traj_offset`` shards a batch over ranks without changing any trajectory.  The reference's
``jax.random`` keys cannot be reproduced without JAX -- same law, different stream.  No CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .functors import Drift, Dispersion, MeasurementFunctor, DriftND, DispersionND, pack_params

__all__ = ['simulate_1d', 'simulate_prey_predator']

_YS_TORCH = {'uint8': torch.uint8, 'int32': torch.int32, 'float64': torch.float64}
_DEFAULT_YS = {'bernoulli_logistic_cubic': 'uint8', 'poisson_softplus': 'int32', 'gaussian': 'float64'}


def _device(device):
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != 'cuda':
        raise ValueError('the simulators run on a CUDA device only (no CPU path)')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    return dev


def _fill(dst, values, what):
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    if values.size > len(dst):
        raise ValueError(f'{what}: at most {_lib.SIM_MAX_COMPONENTS} mixture components')
    for i, v in enumerate(values):
        dst[i] = float(v)


def simulate_1d(drift: Drift, dispersion, dt: float, T: int, init_cond, measurement: MeasurementFunctor, B: int,
                seed: int, integration_steps: int = 100, order: int = 3, scheme: str = 'tme', traj_offset: int = 0,
                device=None, return_xs: bool = False, ys_dtype: str = None):
    """Simulate ``B`` trajectories and their measurements on the device.

    ``drift`` / ``dispersion`` / ``measurement`` are the functor handles of ``mfs_b200.one_dim.ss_models``;
    ``init_cond`` is a ``GaussianSum1D`` (its ``means, variances, weights`` define x0).  ``scheme='tme'`` is the
    reference's simulator (``order`` = TME order, 100 sub-steps by default, ``mfs/one_dim/ss_models.py:49-54``),
    ``'benes_exact'`` draws from the exact Benes transition law.  Per-trajectory parameters (arrays of length B in the
    handles) are allowed.  Returns ``(x0 (B,), xs (B, T) or None, ys (B, T))`` as tensors on ``device``.
    """
    if not isinstance(drift, Drift) or not isinstance(measurement, MeasurementFunctor):
        raise TypeError('drift / measurement must be functor handles (mfs_b200.functors); there is no CPU path')
    if scheme not in _lib.SIM_SCHEME:
        raise ValueError(f'unknown scheme {scheme!r}; known: {sorted(_lib.SIM_SCHEME)}')
    dev = _device(device)
    b = float(dispersion.value) if isinstance(dispersion, Dispersion) else float(dispersion)
    ys_dtype = ys_dtype or _DEFAULT_YS[measurement.name]
    if ys_dtype not in _YS_TORCH:
        raise ValueError(f'ys_dtype must be one of {sorted(_YS_TORCH)}')
    tp, tstride = pack_params(drift.params, (B,))
    mp, mstride = pack_params(measurement.params, (B,))
    a = _lib.Simulate1dArgs()
    a.abi_version = _lib.ABI_VERSION
    a.scheme = _lib.SIM_SCHEME[scheme]
    a.tme_order, a.integration_steps = int(order), int(integration_steps)
    a.drift_id, a.meas_id = _lib.DRIFT[drift.name], _lib.MEAS[measurement.name]
    a.ys_dtype = _lib.YS_DTYPE[ys_dtype]
    a.n_components = int(np.size(init_cond.weights))
    a.B, a.T, a.dt, a.dispersion = int(B), int(T), float(dt), b
    _fill(a.init_means, init_cond.means, 'means')
    _fill(a.init_variances, init_cond.variances, 'variances')
    _fill(a.init_weights, init_cond.weights, 'weights')
    a.seed, a.traj_offset = int(seed) & (2 ** 64 - 1), int(traj_offset)
    with torch.cuda.device(dev):
        tp_d, mp_d = torch.from_numpy(tp).to(dev), torch.from_numpy(mp).to(dev)
        x0 = torch.empty(B, dtype=torch.float64, device=dev)
        ys = torch.empty((B, T), dtype=_YS_TORCH[ys_dtype], device=dev)
        xs = torch.empty((B, T), dtype=torch.float64, device=dev) if return_xs else None
        a.trans_params, a.trans_param_stride = tp_d.data_ptr(), tstride
        a.meas_params, a.meas_param_stride = mp_d.data_ptr(), mstride
        a.ys_out, a.ys_stride_b, a.ys_stride_t = ys.data_ptr(), T, 1
        a.xs_out, a.xs_stride_b, a.xs_stride_t = (xs.data_ptr() if return_xs else None), T, 1
        a.x0_out = x0.data_ptr()
        _lib.check(_lib.lib().mfs_simulate_1d(ctypes.byref(a), torch.cuda.current_stream(dev).cuda_stream))
        tp_d.record_stream(torch.cuda.current_stream(dev))
        mp_d.record_stream(torch.cuda.current_stream(dev))
    return x0, xs, ys


def simulate_prey_predator(drift: DriftND, dispersion: DispersionND, dt: float, T: int, gs,
                           measurement: MeasurementFunctor, B: int, seed: int, integration_steps: int = 100,
                           obs_dim: int = 0, traj_offset: int = 0, device=None, return_xs: bool = False):
    """``prey_predator(...).simulate`` (``mfs/multi_dims/ss_models.py:76-93``) for ``B`` trajectories on the device:
    x0 ~ ``gs`` (a ``GaussianSumND``), Milstein sub-steps, y ~ Bernoulli(emission(x[obs_dim])).
    Returns ``(x0 (B, 2), xs (B, T, 2) or None, ys (B, T) uint8)``."""
    if not isinstance(drift, DriftND) or not isinstance(dispersion, DispersionND):
        raise TypeError('drift / dispersion must be the Lotka--Volterra functor handles; there is no CPU path')
    if measurement.name != 'bernoulli_logistic_cubic':
        raise ValueError('the prey--predator measurement is bernoulli_logistic_cubic (ss_models.py:63-67)')
    dev = _device(device)
    tp, tstride = pack_params(tuple(drift.params) + tuple(dispersion.params), (B,), width=8)
    mp, mstride = pack_params(measurement.params, (B,))
    a = _lib.SimulateLvArgs()
    a.abi_version = _lib.ABI_VERSION
    a.integration_steps, a.obs_dim = int(integration_steps), int(obs_dim)
    means, covs, weights = (np.asarray(v, dtype=np.float64) for v in (gs.means, gs.covs, gs.weights))
    K = weights.size
    if K > _lib.SIM_MAX_COMPONENTS or means.shape != (K, 2) or covs.shape != (K, 2, 2):
        raise ValueError('gs must be a 2-D Gaussian sum with at most 8 components')
    a.n_components = K
    for k in range(K):
        a.init_weights[k] = float(weights[k])
        for i in range(2):
            a.init_means[k][i] = float(means[k, i])
        for i in range(4):
            a.init_covs[k][i] = float(covs[k].reshape(-1)[i])
    a.B, a.T, a.dt = int(B), int(T), float(dt)
    a.seed, a.traj_offset = int(seed) & (2 ** 64 - 1), int(traj_offset)
    with torch.cuda.device(dev):
        tp_d, mp_d = torch.from_numpy(tp).to(dev), torch.from_numpy(mp).to(dev)
        x0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
        ys = torch.empty((B, T), dtype=torch.uint8, device=dev)
        xs = torch.empty((B, T, 2), dtype=torch.float64, device=dev) if return_xs else None
        a.trans_params, a.trans_param_stride = tp_d.data_ptr(), tstride
        a.meas_params, a.meas_param_stride = mp_d.data_ptr(), mstride
        a.ys_out, a.xs_out, a.x0_out = ys.data_ptr(), (xs.data_ptr() if return_xs else None), x0.data_ptr()
        _lib.check(_lib.lib().mfs_simulate_lv(ctypes.byref(a), torch.cuda.current_stream(dev).cuda_stream))
        tp_d.record_stream(torch.cuda.current_stream(dev))
        mp_d.record_stream(torch.cuda.current_stream(dev))
    return x0, xs, ys

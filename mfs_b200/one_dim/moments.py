"""Host-side mirror of ``mfs/one_dim/moments.py``: transition-moment *factories* returning device-functor handles,
plus the small moment conversions the reference's callers and tests use (NumPy, host-side, O(N^2) per vector)."""
import math

import numpy as np
import scipy.linalg

from ..functors import Drift, Dispersion, TransitionSpec, TransitionFunctor

__all__ = ['sde_cond_moments_tme', 'sde_cond_moments_tme_normal', 'sde_cond_moments_euler',
           'sde_cond_moments_normal_affine', 'raw_moment_of_normal', 'raw_moment_of_standard_normal',
           'central_moment_of_normal', 'raw_to_central', 'central_to_raw', 'raw_to_scaled', 'scaled_to_central',
           'characteristic_fn']

_ROLES = ('raw', 'central', 'scaled', 'mean', 'mean_var')


def _dispersion_value(dispersion) -> float:
    if isinstance(dispersion, Dispersion):
        return float(dispersion.value)
    if isinstance(dispersion, (int, float)):
        return float(dispersion)
    raise TypeError('dispersion must be a Dispersion handle or a float: only constant dispersions are registered '
                    '(every 1D model of the reference has b = const).')


def _family(spec: TransitionSpec):
    return tuple(TransitionFunctor(spec, role) for role in _ROLES)


def _check_drift(drift):
    if not isinstance(drift, Drift):
        raise TypeError(f'drift must be a registered Drift handle (benes_drift(), well_drift(theta1), '
                        f'linear_drift(a)), got {type(drift).__name__}: Python callables cannot run inside the CUDA '
                        f'kernel and there is no CPU fallback.')


def sde_cond_moments_tme(drift: Drift, dispersion, dt: float, tme_order: int):
    """Mirror of ``mfs/one_dim/moments.py:141-179``.  Returns the 5-tuple (raw, central, scaled, mean, mean_var)."""
    _check_drift(drift)
    if tme_order not in (1, 2, 3):
        raise ValueError('tme_order must be 1, 2 or 3')
    return _family(TransitionSpec('tme', drift, _dispersion_value(dispersion), float(dt), int(tme_order)))


def sde_cond_moments_tme_normal(drift: Drift, dispersion, dt: float, tme_order: int, N: int = None):
    """Mirror of ``mfs/one_dim/moments.py:182-219`` (``N`` is implied by the moment vectors here)."""
    _check_drift(drift)
    if tme_order not in (1, 2, 3):
        raise ValueError('tme_order must be 1, 2 or 3')
    return _family(TransitionSpec('tme_normal', drift, _dispersion_value(dispersion), float(dt), int(tme_order)))


def sde_cond_moments_euler(drift: Drift, dispersion, dt: float, N: int = None):
    """Mirror of ``mfs/one_dim/moments.py:222-255``."""
    _check_drift(drift)
    return _family(TransitionSpec('euler', drift, _dispersion_value(dispersion), float(dt), 1))


def sde_cond_moments_normal_affine(F, Sigma, N: int = None):
    """Exact Normal transition N(F x, Sigma) -- the hand-written OU closures of
    ``dardel/convergence/convergence_mf.py:86-107``.  ``F``/``Sigma`` scalar or per-filter arrays."""
    return _family(TransitionSpec('normal_affine', Drift('linear', (0.,)), 1., 1., 1, (F, Sigma)))


# ---- host-side helpers (mfs/one_dim/moments.py:31-138) --------------------------------------------------------------
def central_moment_of_normal(variance, p: int):
    if p % 2 == 0:
        return math.sqrt(variance) ** p * math.prod(range(p - 1, 0, -2))
    return 0.


def raw_moment_of_standard_normal(p: int) -> float:
    if p % 2 == 0:
        return math.factorial(p) / (2 ** (p / 2) * math.factorial(p // 2))
    return 0.


def raw_moment_of_normal(mean, variance, p: int):
    return sum(math.comb(p, m) * mean ** m * variance ** ((p - m) / 2) * raw_moment_of_standard_normal(p - m)
               for m in range(p + 1))


def _pascal(s):
    return scipy.linalg.pascal(s, kind='lower', exact=True).astype(np.float64)


def raw_to_central(rms):
    rms = np.asarray(rms, dtype=np.float64)
    s = rms.shape[-1]
    bn = _pascal(s)
    n, j = np.meshgrid(np.arange(s), np.arange(s), indexing='ij')
    with np.errstate(all='ignore'):
        coeff = np.where(n >= j, bn * (-1.) ** np.maximum(n - j, 0), 0.)
        powers = rms[..., 1, None, None] ** np.maximum(n - j, 0)
    return np.sum(coeff * rms[..., None, :] * powers, axis=-1)


def central_to_raw(cms, mean):
    cms = np.asarray(cms, dtype=np.float64)
    mean = np.asarray(mean, dtype=np.float64)
    s = cms.shape[-1]
    bn = _pascal(s)
    n, j = np.meshgrid(np.arange(s), np.arange(s), indexing='ij')
    coeff = np.where(n >= j, bn, 0.)
    powers = mean[..., None, None] ** np.maximum(n - j, 0)
    return np.sum(coeff * cms[..., None, :] * powers, axis=-1)


def raw_to_scaled(rms, scale=None):
    rms = np.asarray(rms, dtype=np.float64)
    if scale is None:
        scale = np.sqrt(rms[..., 2] - rms[..., 1] ** 2)
    scale = np.asarray(scale, dtype=np.float64)
    return raw_to_central(rms) / scale[..., None] ** np.arange(rms.shape[-1])


def scaled_to_central(sms, scale):
    sms = np.asarray(sms, dtype=np.float64)
    return sms * np.asarray(scale, dtype=np.float64)[..., None] ** np.arange(sms.shape[-1])


def characteristic_fn(z, ms, mean=0., scale=1.):
    """Characteristic function by moments, mirror of ``mfs/one_dim/moments.py:309-337``:
    ``E[exp(i z X)] ~ sum_n w_n exp(i z x_n)`` with ``(w, x) = moment_quadrature(ms, mean, scale)``.

    The reference vmaps this scalar function over the z-grid and the moment history
    (``dardel/benes_bernoulli/post_processing_mf.py:37-60``); here ``z`` may be a scalar or ``(m,)`` and ``ms``
    ``(2n,)`` or ``(..., 2n)`` (``mean`` / ``scale`` broadcast against the batch axes): one kernel launch, result
    ``(..., m)`` complex128 (``(...)`` for scalar ``z``).  torch CUDA ``ms`` -> torch CUDA result, NumPy in -> NumPy out.
    """
    import ctypes
    import torch
    from .. import _lib
    is_np = not (isinstance(ms, torch.Tensor) and ms.is_cuda)
    dev = torch.device('cuda', torch.cuda.current_device()) if is_np else ms.device
    ms_t = torch.as_tensor(np.asarray(ms, dtype=np.float64) if is_np else ms, dtype=torch.float64, device=dev)
    n = math.floor(ms_t.shape[-1] / 2)
    batch_shape = tuple(ms_t.shape[:-1])
    B = int(np.prod(batch_shape)) if batch_shape else 1
    ms_c = ms_t[..., :2 * n].reshape(B, 2 * n).contiguous()
    scalar_z = np.ndim(z) == 0 and not isinstance(z, torch.Tensor)
    zs = torch.as_tensor(np.atleast_1d(np.asarray(z, dtype=np.float64)) if not isinstance(z, torch.Tensor) else z,
                         dtype=torch.float64, device=dev).reshape(-1).contiguous()
    m = int(zs.shape[0])

    def aux(v, default):
        if np.ndim(v) == 0 and not isinstance(v, torch.Tensor) and float(v) == default:
            return None
        t = torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                            dtype=torch.float64, device=dev)
        return t.expand(batch_shape).reshape(B).contiguous()

    mean_t, scale_t = aux(mean, 0.), aux(scale, 1.)
    out = torch.empty((B, m), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().mfs_characteristic_fn_1d(
            n, B, m, ms_c.data_ptr(), None if mean_t is None else mean_t.data_ptr(),
            None if scale_t is None else scale_t.data_ptr(), zs.data_ptr(), out.data_ptr(), ctypes.c_void_p(stream)))
    out = out.reshape(batch_shape + (m,))
    if scalar_z:
        out = out[..., 0]
    return out.cpu().numpy() if is_np else out

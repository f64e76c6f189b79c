"""Host-side mirror of ``mfs/multi_dims/filtering.py`` (d = 2): same names and positional signatures; the scan runs in
one CUDA kernel launch (one warp per filter, ``mfs_b200/csrc/filter_nd.cuh``).

``ys`` may carry leading batch axes ``(..., T)``; initial moments ``(z,)`` or ``(..., z)``; ``mean0`` ``(d,)`` or
``(..., d)``.  A torch CUDA tensor for ``ys`` keeps everything on the device (CUDA tensors out); NumPy input is copied to
the current CUDA device and the results come back as NumPy arrays.  No CPU compute path.
"""
import ctypes

import numpy as np

from .. import _lib
from ..functors import MeasurementFunctor, pack_params
from .moments import TransitionFunctorND
from .multi_indices import generate_graded_lexico_multi_indices, gram_and_hankel_indices_graded_lexico

__all__ = ['moment_filter_nd_rms', 'moment_filter_nd_cms']


def _check(fn_and_flag, role):
    if not (isinstance(fn_and_flag, (tuple, list)) and len(fn_and_flag) == 2):
        raise TypeError("pass the conditional-moment functor as the reference does: (functor, 'index' | 'multi-index')")
    fn, flag = fn_and_flag
    if flag not in ('index', 'multi-index'):
        raise ValueError("signature flag must be 'index' or 'multi-index' (mfs/multi_dims/filtering.py:120-123)")
    if not isinstance(fn, TransitionFunctorND):
        raise TypeError('the conditional-moment callable must be a TransitionFunctorND handle from '
                        'mfs_b200.multi_dims.moments.sde_cond_moments_*; Python callables cannot run in the kernel.')
    if fn.role != role:
        raise ValueError(f'expected the {role!r} member of the factory tuple, got {fn.role!r}')
    want = 'multi-index' if fn.spec.family == 'tme' else 'index'      # moments.py:296-337 vs :441-459
    if flag != want:
        raise ValueError(f"the {fn.spec.family!r} factory has the {want!r} signature (got flag {flag!r})")
    return fn.spec


def _run(mode, spec, meas, ys, moments_partial_order, ms0, mean0, stable, history, return_status):
    import torch
    if not isinstance(meas, MeasurementFunctor) or meas.name != 'bernoulli_logistic_cubic':
        raise TypeError('measurement_cond_pdf must be the bernoulli_logistic_cubic MeasurementFunctor handle')
    multi_indices, inds = moments_partial_order
    multi_indices, inds = np.asarray(multi_indices), np.asarray(inds)
    d = multi_indices.shape[-1]
    if d != 2:
        raise _lib.MfsError(f'only d = 2 is implemented (got d = {d})')
    z = multi_indices.shape[0]
    ms0 = np.asarray(ms0.cpu() if isinstance(ms0, torch.Tensor) else ms0, dtype=np.float64)
    if z != ms0.shape[-1]:
        raise ValueError(f'The size of multi_indices {z} must match that of the initial moments {ms0.shape[-1]}.')
    s = inds.shape[1]
    N = int(round((np.sqrt(8 * s + 1) - 1) / 2))
    # the kernel uses the closed-form graded-lex position (a+b)(a+b+1)/2 + a: insist the tables are in that order
    if not (np.array_equal(multi_indices, generate_graded_lexico_multi_indices(2, 2 * N - 1, 0))
            and np.array_equal(inds, gram_and_hankel_indices_graded_lexico(N, 2))):
        raise ValueError('moments_partial_order must be (generate_graded_lexico_multi_indices(2, 2N-1), '
                         'gram_and_hankel_indices_graded_lexico(N, 2))')
    is_np = not (isinstance(ys, torch.Tensor) and ys.is_cuda)
    dev = torch.device('cuda', torch.cuda.current_device()) if is_np else ys.device
    ys_t = torch.as_tensor(np.asarray(ys).astype(np.uint8) if is_np else ys, device=dev)
    if ys_t.dtype == torch.bool:
        ys_t = ys_t.view(torch.uint8)
    if ys_t.dtype != torch.uint8:
        ys_t = ys_t.to(torch.uint8)
    batch_shape = tuple(ys_t.shape[:-1])
    T = int(ys_t.shape[-1])
    B = int(np.prod(batch_shape)) if batch_shape else 1
    ys_c = ys_t.reshape(B, T).contiguous()

    def table(arr, trailing):
        arr = np.asarray(arr.cpu() if isinstance(arr, torch.Tensor) else arr, dtype=np.float64)
        if arr.shape == trailing:
            return torch.from_numpy(np.ascontiguousarray(arr.reshape((1,) + trailing))).to(dev), 0
        full = np.ascontiguousarray(np.broadcast_to(arr, batch_shape + trailing).reshape((B,) + trailing))
        return torch.from_numpy(full).to(dev), int(np.prod(trailing))

    ms0_t, ms0_stride = table(ms0, (z,))
    tprm_np, tstride = pack_params(spec.packed_params(), batch_shape, width=8)
    mprm_np, mstride = pack_params(meas.params, batch_shape)
    tprm, mprm = torch.from_numpy(tprm_np).to(dev), torch.from_numpy(mprm_np).to(dev)
    inds_t = torch.from_numpy(np.ascontiguousarray(inds.astype(np.int32))).to(dev)

    a = _lib.FilterNdArgs()
    a.abi_version, a.mode, a.N, a.d, a.B, a.T = _lib.ABI_VERSION, _lib.MODE[mode], N, 2, B, T
    a.trans_id, a.tme_order = _lib.TRANS[spec.family], spec.order
    a.meas_id, a.obs_dim, a.dt = _lib.MEAS[meas.name], 0, spec.dt
    a.trans_params, a.trans_param_stride = tprm.data_ptr(), tstride
    a.meas_params, a.meas_param_stride = mprm.data_ptr(), mstride
    a.ms0, a.ms0_stride = ms0_t.data_ptr(), ms0_stride
    keep = [ys_c, ms0_t, tprm, mprm, inds_t]
    if mode == 'central':
        mean0_t, mean0_stride = table(mean0, (2,))
        a.mean0, a.mean0_stride = mean0_t.data_ptr(), mean0_stride
        keep.append(mean0_t)
    a.ys, a.inds = ys_c.data_ptr(), inds_t.data_ptr()
    a.out_mode = _lib.OUT_MODE[history]
    a.stable = int(bool(stable))
    f64 = dict(dtype=torch.float64, device=dev)
    nell = torch.empty(B, **f64)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    ms_out = mean_out = None
    tail = (T,) if history in ('full', 'meanvar') else ()
    width = 5 if history == 'meanvar' else z
    if history != 'none':
        ms_out = torch.empty((B,) + tail + (width,), **f64)
        a.ms_out = ms_out.data_ptr()
        if mode == 'central' and history != 'meanvar':
            mean_out = torch.empty((B,) + tail + (2,), **f64)
            a.mean_out = mean_out.data_ptr()
    a.nell_out, a.status_out = nell.data_ptr(), status.data_ptr()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().mfs_filter_nd(ctypes.byref(a), ctypes.c_void_p(stream)))
    for t in keep:
        t.record_stream(torch.cuda.current_stream(dev))

    def fin(t, tl):
        if t is None:
            return None
        t = t.reshape(batch_shape + tl)
        return t.cpu().numpy() if is_np else t

    return fin(ms_out, tail + (width,)), fin(mean_out, tail + (2,)), fin(nell, ()), fin(status, ())


def moment_filter_nd_rms(state_cond_raw_moments, measurement_cond_pdf, ys, moments_partial_order, rms0,
                         stable: bool = False, *, history: str = 'full', return_status: bool = False):
    """Mirror of ``mfs/multi_dims/filtering.py:283-344``.  Returns ``(rmss (..., T, z), nell (...))``.
    ``history='meanvar'``: ``rmss`` is ``(..., T, 5)`` = (E x1, E x2, Var x1, Cov, Var x2) per step -- what the reference's
    consumer keeps (``dardel/prey_predator/mf.py:84-88``); ``'last'`` / ``'none'``: final moments only / nell only."""
    spec = _check(state_cond_raw_moments, 'raw')
    ms, _, nell, status = _run('raw', spec, measurement_cond_pdf, ys, moments_partial_order, rms0, None, stable,
                               history, return_status)
    return (ms, nell, status) if return_status else (ms, nell)


def moment_filter_nd_cms(state_cond_central_moments, state_cond_mean, measurement_cond_pdf, ys, moments_partial_order,
                         cms0, mean0, stable: bool = False, *, history: str = 'full', return_status: bool = False):
    """Mirror of ``mfs/multi_dims/filtering.py:210-280``.  Returns ``(cmss (..., T, z), means (..., T, d), nell)``.
    ``history='meanvar'``: ``cmss`` is ``(..., T, 5)`` = (E x1, E x2, Var x1, Cov, Var x2) per step and ``means`` is None."""
    spec = _check(state_cond_central_moments, 'central')
    if not isinstance(state_cond_mean, TransitionFunctorND) or state_cond_mean.role != 'mean' \
            or state_cond_mean.spec is not spec:
        raise ValueError('state_cond_mean must be the "mean" member of the same factory call')
    ms, means, nell, status = _run('central', spec, measurement_cond_pdf, ys, moments_partial_order, cms0, mean0,
                                   stable, history, return_status)
    return (ms, means, nell, status) if return_status else (ms, means, nell)

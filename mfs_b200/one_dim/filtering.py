"""Host-side mirror of ``mfs/one_dim/filtering.py``: same names, positional signatures, return tuples and warnings;
the scan runs in one CUDA kernel launch for the whole batch (``mfs_b200/csrc/filter1d.cuh``).

Extension over the reference: leading batch axes.  ``ys`` may be ``(..., T)`` and the initial moments ``(2N,)`` or
``(..., 2N)``; functor parameters may be per-filter arrays.  Every filter of the batch is independent.

Where the data lives decides the path (there is no CPU compute path):
  * ``ys`` is a ``torch`` CUDA tensor  -> device pointers are handed to ``mfs_filter_1d`` on torch's current stream,
    outputs are CUDA tensors (asynchronous w.r.t. the host);
  * ``ys`` is a NumPy array / CPU tensor -> ``mfs_filter_1d_host`` stages chunks H2D -> kernel -> D2H on ``device``
    and returns NumPy arrays.
"""
import ctypes
import warnings
from typing import Optional

import numpy as np

from .. import _lib
from ..functors import TransitionFunctor, MeasurementFunctor, pack_params

__all__ = ['moment_filter_rms', 'moment_filter_cms', 'moment_filter_scms']


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith('torch') and getattr(x, 'is_cuda', False)


def _check_transition(fn, role: str, argname: str) -> TransitionFunctor:
    if not isinstance(fn, TransitionFunctor):
        raise TypeError(f'{argname} must be a TransitionFunctor handle produced by mfs_b200.one_dim.moments.'
                        f'sde_cond_moments_*; Python callables cannot run inside the CUDA kernel (no CPU fallback).')
    if fn.role != role:
        raise ValueError(f'{argname} must be the {role!r} member of the factory tuple, got {fn.role!r}')
    return fn


def _check_measurement(fn) -> MeasurementFunctor:
    if not isinstance(fn, MeasurementFunctor):
        raise TypeError('measurement_cond_pdf must be a MeasurementFunctor handle (bernoulli_logistic_cubic, '
                        'poisson_softplus, gaussian); Python callables cannot run inside the CUDA kernel.')
    return fn


def _ys_dtype_code(dtype_name: str) -> int:
    name = dtype_name.replace('torch.', '')
    if name not in _lib.YS_DTYPE:
        raise TypeError(f'ys dtype {dtype_name} not supported (uint8 / bool / int32 / float64)')
    return _lib.YS_DTYPE[name]


def _run(mode: str, spec, meas: MeasurementFunctor, ms0, mean0, scale0, ys, stable: bool, history: str,
         device: Optional[int], return_status: bool, chunk_filters: int, out: Optional[dict] = None,
         recompute_predict_quadrature: bool = False, segment_steps: Optional[int] = None,
         grad_ids: Optional[list] = None, carry=None, return_carry: bool = False, t_offset: int = 0):
    if history not in _lib.OUT_MODE:
        raise ValueError(f"history must be one of {sorted(_lib.OUT_MODE)}")
    on_device = _is_torch_cuda(ys)
    if grad_ids is not None and not on_device:
        raise ValueError('value_and_grad takes ys as a CUDA tensor (device path only)')
    if (carry is not None or return_carry) and not on_device:
        raise ValueError('carry / return_carry (time-chunked execution) take ys as a CUDA tensor (device path only)')
    if not on_device and type(ys).__module__.startswith('torch'):
        ys = ys.numpy()
    if not on_device:
        ys = np.asarray(ys)
        if ys.dtype == np.bool_:
            ys = ys.view(np.uint8)
        elif ys.dtype == np.int64:
            ys = ys.astype(np.int32)
        elif ys.dtype in (np.float32, np.float16):
            ys = ys.astype(np.float64)
    if ys.ndim < 1:
        raise ValueError('ys must have shape (..., T)')
    T = int(ys.shape[-1])
    # Parameter grid over shared records (theta grids): when the functor parameters / initial moments carry LEADING
    # batch axes that ys does not have -- e.g. ys (n_traj, T) with parameters shaped (n_theta, 1) -- the batch is the
    # (n_theta, n_traj) grid and the kernel reads record j and parameter row g for filter (g, j); nothing is
    # materialised per filter (mfs_filter1d_args.grid_records).
    rec_shape = tuple(ys.shape[:-1])
    extra = [np.shape(p) for p in tuple(spec.packed_params()) + tuple(meas.params) if np.ndim(p) > 0]
    extra.append(np.shape(ms0)[:-1])
    if mode != 'raw':
        extra.append(np.shape(mean0))
    if mode == 'scaled':
        extra.append(np.shape(scale0))
    full_shape = tuple(np.broadcast_shapes(rec_shape, *extra))
    grid_records = 0
    if full_shape != rec_shape:
        k = len(rec_shape)
        lead = full_shape[:len(full_shape) - k]
        aligned = k == 0 or full_shape[-k:] == rec_shape
        const_along_records = all(all(d == 1 for d in sh[max(0, len(sh) - k):]) if k else True for sh in extra)
        if aligned and const_along_records and on_device and grad_ids is None and carry is None and not return_carry:
            grid_records = int(np.prod(rec_shape)) if rec_shape else 1
            grid_shape = lead
        else:   # general broadcasting: one record per filter, materialised
            if on_device:
                ys = ys.expand(full_shape + (T,))
            else:
                ys = np.broadcast_to(ys, full_shape + (T,))
    batch_shape = full_shape
    B = int(np.prod(batch_shape)) if batch_shape else 1

    ms0_np = np.asarray(ms0.cpu() if _is_torch_cuda(ms0) else ms0, dtype=np.float64)
    num_moments = ms0_np.shape[-1]
    if num_moments % 2 != 0:
        # mfs/one_dim/filtering.py:65-66 warns and carries the extra top moment along without ever using it in a
        # quadrature; here the kernel works on the 2N moments the quadrature uses and the returned history has 2N
        # columns (the reference's has 2N + 1).
        warnings.warn(f'The order of moments {num_moments - 1} is not odd. The top moment is dropped: the filter '
                      f'runs on the first {num_moments - 1} moments and returns {num_moments - 1} columns.')
    N = num_moments // 2
    if N < 2 or N > _lib.MAX_N:
        raise ValueError(f'number of moments {num_moments} outside [4, {2 * _lib.MAX_N}]')
    M = 2 * N

    # shape of the per-filter input tables: the whole batch, or only the grid's leading axes
    table_shape = grid_shape if grid_records else batch_shape
    n_rows = int(np.prod(table_shape)) if table_shape else 1
    rec_ones = (1,) * len(rec_shape) if grid_records else ()

    def on_table(arr, trailing=()):
        """broadcast a per-filter input to the table's batch axes (grid mode: its size-1 record axes are dropped)"""
        arr = np.asarray(arr.cpu() if _is_torch_cuda(arr) else arr, dtype=np.float64)
        return np.broadcast_to(arr, table_shape + rec_ones + trailing).reshape(table_shape + trailing)

    def per_filter(arr, trailing):
        """-> (contiguous float64 array (rows, *trailing), stride in elements)"""
        arr = np.asarray(arr.cpu() if _is_torch_cuda(arr) else arr, dtype=np.float64)
        if arr.shape == trailing:
            return np.ascontiguousarray(arr.reshape((1,) + trailing)), 0
        full = on_table(arr, trailing).reshape((n_rows,) + trailing)
        return np.ascontiguousarray(full), int(np.prod(trailing)) if trailing else 1

    ms0_tab, ms0_stride = per_filter(ms0_np[..., :M], (M,))
    mean0_tab = scale0_tab = None
    mean0_stride = scale0_stride = 0
    if mode != 'raw':
        mean0_tab, mean0_stride = per_filter(mean0, ())
    if mode == 'scaled':
        scale0_tab, scale0_stride = per_filter(scale0, ())
    tprm, tstride = pack_params(tuple(on_table(p) if np.ndim(p) > 0 else p for p in spec.packed_params()), table_shape)
    mprm, mstride = pack_params(tuple(on_table(p) if np.ndim(p) > 0 else p for p in meas.params), table_shape)

    a = _lib.Filter1dArgs()
    a.abi_version, a.mode, a.N, a.stable = _lib.ABI_VERSION, _lib.MODE[mode], N, int(bool(stable))
    a.B, a.T = B, T
    a.trans_id, a.drift_id = _lib.TRANS[spec.family], _lib.DRIFT[spec.drift.name]
    a.tme_order, a.meas_id = spec.order, _lib.MEAS[meas.name]
    a.dt, a.dispersion = spec.dt, spec.dispersion
    a.trans_param_stride, a.meas_param_stride = tstride, mstride
    a.ms0_stride, a.mean0_stride, a.scale0_stride = ms0_stride, mean0_stride, scale0_stride
    a.out_mode = _lib.OUT_MODE[history]
    a.flags = _lib.FLAG_RECOMPUTE_PREDICT_QUADRATURE if recompute_predict_quadrature else 0
    a.ys_stride_b, a.ys_stride_t = T, 1
    a.grid_records = grid_records

    L = _lib.lib()
    if on_device:
        import torch
        dev = ys.device
        if ys.dim() == 2 and not ys.is_contiguous() and min(ys.stride()) > 0 and ys.dtype != torch.bool:
            # a strided 2-D view (e.g. the transpose of a time-major (T, B) buffer) is read in place through the ABI's
            # element strides: with stride_b = 1 consecutive lanes read consecutive measurements
            ys_c = ys
            a.ys_stride_b, a.ys_stride_t = int(ys.stride(0)), int(ys.stride(1))
        else:
            ys_c = ys.reshape(grid_records if grid_records else B, T)
            if ys_c.dtype == torch.bool:
                ys_c = ys_c.view(torch.uint8)
            ys_c = ys_c.contiguous()
        a.ys_dtype = _ys_dtype_code(str(ys_c.dtype))
        keep = [ys_c]

        def to_dev(tab):
            t = torch.from_numpy(tab).to(dev)
            keep.append(t)
            return t.data_ptr()

        a.ys = ys_c.data_ptr()
        a.ms0, a.trans_params, a.meas_params = to_dev(ms0_tab), to_dev(tprm), to_dev(mprm)
        if mean0_tab is not None:
            a.mean0 = to_dev(mean0_tab)
        if scale0_tab is not None:
            a.scale0 = to_dev(scale0_tab)
        f64 = dict(dtype=torch.float64, device=dev)
        out = out or {}

        def alloc(key, shape, dtype=torch.float64):
            t = out.get(key)
            if t is None:
                return torch.empty(shape, dtype=dtype, device=dev)
            if not (t.is_cuda and t.dtype == dtype and t.is_contiguous() and t.numel() == int(np.prod(shape))):
                raise ValueError(f"out[{key!r}] must be a contiguous {dtype} CUDA tensor with {int(np.prod(shape))} elements")
            return t.view(shape)

        nell = alloc('nell', (B,))
        status = alloc('status', (B,), torch.int32)
        ms_out = mean_out = scale_out = None
        if history == 'full':
            ms_out = alloc('ms', (B, T, M))
            a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = T * M, M, T
            aux_shape = (B, T)
        elif history == 'last':
            ms_out = alloc('ms', (B, M))
            a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = M, 0, 1
            aux_shape = (B,)
        elif history == 'meanvar':
            ms_out = alloc('meanvar', (B, T, 2))
            a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = T * 2, 2, T
        if ms_out is not None:
            a.ms_out = ms_out.data_ptr()
            if mode != 'raw' and history != 'meanvar':
                mean_out = alloc('mean', aux_shape)
                a.mean_out = mean_out.data_ptr()
            if mode == 'scaled' and history != 'meanvar':
                scale_out = alloc('scale', aux_shape)
                a.scale_out = scale_out.data_ptr()
        a.nell_out, a.status_out = nell.data_ptr(), status.data_ptr()
        carry_out = None
        if carry is not None:
            if not (carry.is_cuda and carry.dtype == torch.float64 and carry.is_contiguous()
                    and carry.numel() == B * (4 * N + 4)):
                raise ValueError(f'carry must be a contiguous float64 CUDA tensor with {B} x {4 * N + 4} elements')
            keep.append(carry)
            a.carry_in = carry.data_ptr()
        if return_carry:
            carry_out = alloc('carry', (B, 4 * N + 4))
            a.carry_out = carry_out.data_ptr()
        a.t_offset = int(t_offset)
        # long horizons: scratch for the segmented execution with live-filter compaction (include/mfs_b200.h)
        seg = 64 if segment_steps is None else int(segment_steps)
        if seg > 0 and T >= 2 * seg:
            ws_bytes = int(L.mfs_filter_1d_workspace_bytes(N, B, T))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            keep.append(ws)
            a.segment_steps, a.workspace, a.workspace_bytes = seg, ws.data_ptr(), ws_bytes
        grad = None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if grad_ids is None:
                _lib.check(L.mfs_filter_1d(ctypes.byref(a), ctypes.c_void_p(stream)))
            else:
                # forward-mode nell + gradient kernel, MFS_GRAD_MAX_TANGENTS parameters per pass
                W = _lib.GRAD_MAX_TANGENTS
                grad = torch.zeros((B, len(grad_ids)), **f64)
                for k0 in range(0, len(grad_ids), W):
                    ids = list(grad_ids[k0:k0 + W])
                    g = torch.empty((B, W), **f64)
                    keep.append(g)
                    _lib.check(L.mfs_filter_1d_grad(ctypes.byref(a), len(ids), (ctypes.c_int32 * len(ids))(*ids),
                                                    ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(stream)))
                    grad[:, k0:k0 + len(ids)] = g[:, :len(ids)]
        for t in keep:   # inputs must outlive the asynchronous kernel
            t.record_stream(torch.cuda.current_stream(dev))
        reshape = lambda t, tail: None if t is None else t.reshape(batch_shape + tail)
    else:
        ys_c = np.ascontiguousarray(ys.reshape(B, T))
        a.ys_dtype = _ys_dtype_code(str(ys_c.dtype))
        ptr = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
        a.ys = ptr(ys_c)
        a.ms0, a.trans_params, a.meas_params = ptr(ms0_tab), ptr(tprm), ptr(mprm)
        if mean0_tab is not None:
            a.mean0 = ptr(mean0_tab)
        if scale0_tab is not None:
            a.scale0 = ptr(scale0_tab)
        out = out or {}

        def alloc(key, shape, dtype=np.float64):
            arr = out.get(key)
            if arr is None:
                return np.empty(shape, dtype=dtype)
            if type(arr).__module__.startswith('torch'):
                arr = arr.numpy()       # e.g. a pinned CPU tensor: zero-copy view
            if not (arr.dtype == dtype and arr.flags['C_CONTIGUOUS'] and arr.size == int(np.prod(shape))):
                raise ValueError(f"out[{key!r}] must be a C-contiguous {np.dtype(dtype)} array with {int(np.prod(shape))} elements")
            return arr.reshape(shape)

        nell = alloc('nell', (B,))
        status = alloc('status', (B,), np.int32)
        ms_out = mean_out = scale_out = None
        if history == 'full':
            ms_out = alloc('ms', (B, T, M))
            a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = T * M, M, T
            aux_shape = (B, T)
        elif history == 'last':
            ms_out = alloc('ms', (B, M))
            a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = M, 0, 1
            aux_shape = (B,)
        elif history == 'meanvar':
            ms_out = alloc('meanvar', (B, T, 2))
            a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = T * 2, 2, T
        carry_out = None
        if ms_out is not None:
            a.ms_out = ptr(ms_out)
            if mode != 'raw' and history != 'meanvar':
                mean_out = alloc('mean', aux_shape)
                a.mean_out = ptr(mean_out)
            if mode == 'scaled' and history != 'meanvar':
                scale_out = alloc('scale', aux_shape)
                a.scale_out = ptr(scale_out)
        a.nell_out, a.status_out = ptr(nell), ptr(status)
        if device is None:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        a.segment_steps = -1 if (segment_steps is not None and int(segment_steps) <= 0) else int(segment_steps or 0)
        _lib.check(L.mfs_filter_1d_host(ctypes.byref(a), int(device), int(chunk_filters)))
        reshape = lambda t, tail: None if t is None else t.reshape(batch_shape + tail)

    hist_tail = (T,) if history == 'full' else ()
    out = {
        'ms': reshape(ms_out, (T, 2)) if history == 'meanvar' else reshape(ms_out, hist_tail + (M,)),
        'carry': reshape(carry_out, (4 * N + 4,)),
        'mean': reshape(mean_out, hist_tail),
        'scale': reshape(scale_out, hist_tail),
        'nell': reshape(nell, ()),
        'status': reshape(status, ()),
    }
    if grad_ids is not None:
        out['grad'] = grad.reshape(batch_shape + (len(grad_ids),))
    return out


def moment_filter_rms(state_cond_raw_moments, measurement_cond_pdf, rms0, ys, stable: bool = False, *,
                      history: str = 'full', device: Optional[int] = None, return_status: bool = False,
                      chunk_filters: int = 0, out: Optional[dict] = None,
                      recompute_predict_quadrature: bool = False, segment_steps: Optional[int] = None,
                      carry=None, return_carry: bool = False, t_offset: int = 0):
    """Raw-moment filter, mirror of ``mfs/one_dim/filtering.py:32-89``.

    Returns ``(rmss (..., T, 2N), nell (...))`` like the reference.  ``history='last'`` -> ``(..., 2N)``,
    ``'none'`` -> ``None``, ``'meanvar'`` -> ``(..., T, 2)`` holding (mean, variance) of the filtering distribution per
    step -- what the reference's consumers read from the history (``dardel/prey_predator/mf.py:84-88``,
    ``dardel/benes_bernoulli/post_processing_mf.py:41-66``) at 16 bytes per step instead of 16 N.  With
    ``return_status=True`` a further element holds the first failed step per filter (-1: none).  A filter whose
    moment matrix stops being positive definite yields NaN from that step on, exactly like the JAX scan.  ``out`` may
    hold preallocated result buffers under the keys 'ms' ('meanvar'), 'mean', 'scale', 'nell', 'status', 'carry'
    (CUDA tensors on the device path; NumPy arrays or pinned CPU tensors on the host path) to avoid re-allocation.

    ``recompute_predict_quadrature=True`` forces the literal recursion of the reference (a second ``moment_quadrature``
    per step, ``filtering.py:78``).  By default the prediction half-step re-uses the posterior atoms
    ``{x_i, w_i p(y|x_i)/c}`` of the previous update -- the N-point Gauss rule of an N-atom measure is the measure
    itself -- while still testing the Hankel pivots of the posterior moments, so NaN-on-non-PD fires identically.

    Long horizons (``T >= 2 * segment_steps``, default 64) run segment by segment, each segment launching only the
    filters that are still alive, densely re-packed (a diverged filter would otherwise idle its warp until the end of
    the scan); results are identical.  ``segment_steps=0`` runs the whole scan in one launch.

    Time-chunked execution (checkpoint / resume; device path): ``return_carry=True`` appends the filter state after the
    last step (a ``(..., 4N + 4)`` CUDA tensor) to the returned tuple; passing it as ``carry=`` to the next call, with
    the next block of measurements (and ``t_offset`` = number of steps already filtered, used for the reported status),
    continues the scan bit-identically to one long call -- ``rms0`` is then ignored and ``nell`` is the running total.
    """
    fn = _check_transition(state_cond_raw_moments, 'raw', 'state_cond_raw_moments')
    out = _run('raw', fn.spec, _check_measurement(measurement_cond_pdf), rms0, None, None, ys, stable, history,
               device, return_status, chunk_filters, out, recompute_predict_quadrature, segment_steps,
               carry=carry, return_carry=return_carry, t_offset=t_offset)
    res = (out['ms'], out['nell'])
    res = res + (out['status'],) if return_status else res
    return res + (out['carry'],) if return_carry else res


def moment_filter_cms(state_cond_central_moments, state_cond_mean, measurement_cond_pdf, cms0, mean0, ys,
                      stable: bool = False, *, history: str = 'full', device: Optional[int] = None,
                      return_status: bool = False, chunk_filters: int = 0, out: Optional[dict] = None,
                      recompute_predict_quadrature: bool = False, segment_steps: Optional[int] = None,
                      carry=None, return_carry: bool = False, t_offset: int = 0):
    """Central-moment filter, mirror of ``mfs/one_dim/filtering.py:92-161``.  Returns ``(cmss, means, nell)``
    (``history='meanvar'``: the first element is the ``(..., T, 2)`` (mean, variance) history and ``means`` is None).
    Keyword extensions as in :func:`moment_filter_rms`."""
    fn = _check_transition(state_cond_central_moments, 'central', 'state_cond_central_moments')
    fm = _check_transition(state_cond_mean, 'mean', 'state_cond_mean')
    if fm.spec is not fn.spec:
        raise ValueError('state_cond_central_moments and state_cond_mean must come from the same factory call')
    out = _run('central', fn.spec, _check_measurement(measurement_cond_pdf), cms0, mean0, None, ys, stable, history,
               device, return_status, chunk_filters, out, recompute_predict_quadrature, segment_steps,
               carry=carry, return_carry=return_carry, t_offset=t_offset)
    res = (out['ms'], out['mean'], out['nell'])
    res = res + (out['status'],) if return_status else res
    return res + (out['carry'],) if return_carry else res


def moment_filter_scms(state_cond_scaled_central_moments, state_cond_mean_var, measurement_cond_pdf, scms0, mean0,
                       scale0, ys, stable: bool = False, *, history: str = 'full', device: Optional[int] = None,
                       return_status: bool = False, chunk_filters: int = 0, out: Optional[dict] = None,
                       recompute_predict_quadrature: bool = False, segment_steps: Optional[int] = None,
                       carry=None, return_carry: bool = False, t_offset: int = 0):
    """Scaled-central-moment filter, mirror of ``mfs/one_dim/filtering.py:164-240``.
    Returns ``(scmss, means, scales, nell)``.  With the Normal-approximation factories (``sde_cond_moments_tme_normal``,
    ``sde_cond_moments_euler``) the scaled transition moments are the reference's literally: every order is divided by
    ``prod(scale ** arange(2N))`` (``mfs/one_dim/moments.py:205, 243``), not by ``scale ** order``.
    Keyword extensions as in :func:`moment_filter_rms`."""
    fn = _check_transition(state_cond_scaled_central_moments, 'scaled', 'state_cond_scaled_central_moments')
    fm = _check_transition(state_cond_mean_var, 'mean_var', 'state_cond_mean_var')
    if fm.spec is not fn.spec:
        raise ValueError('state_cond_scaled_central_moments and state_cond_mean_var must come from the same factory')
    out = _run('scaled', fn.spec, _check_measurement(measurement_cond_pdf), scms0, mean0, scale0, ys, stable,
               history, device, return_status, chunk_filters, out, recompute_predict_quadrature, segment_steps,
               carry=carry, return_carry=return_carry, t_offset=t_offset)
    res = (out['ms'], out['mean'], out['scale'], out['nell'])
    res = res + (out['status'],) if return_status else res
    return res + (out['carry'],) if return_carry else res

"""Host-side mirror of ``mfs/classical_filters_smoothers/brute_force.py``: same name, positional signature and
``pred_method`` strings; the work runs in ``mfs_b200/csrc/brute_force.cu`` (FP64 tensor-core GEMM per integration
sub-step, row-wise Bayes update) through the C ABI ``mfs_brute_force``.

Extension over the reference: ``ys`` may be ``(B, T)`` and ``init_ps`` ``(n,)`` or ``(B, n)`` -- B measurement records
filtered on ONE shared grid ``xs`` (that is what makes the transition-density x density product a dense contraction);
or ``xs`` ``(B, n)``: one grid per record, the call shape of ``dardel/benes_bernoulli/brute_force.py:73-76``.
``drift`` / ``dispersion`` / ``measurement_cond_pdf`` are functor handles (``mfs_b200.functors``); a Python callable
raises ``TypeError`` -- there is no CPU path.
"""
import ctypes

import numpy as np

from .. import _lib
from ..functors import Drift, Dispersion, MeasurementFunctor, pack_params

__all__ = ['brute_force_filter']


def brute_force_filter(drift, dispersion, measurement_cond_pdf, init_ps, xs, ys, dt, integration_steps: int = 1,
                       pred_method: str = 'chapman-tme-2', history: str = 'full', return_nell: bool = False,
                       device=None, power_operator: bool = False):
    """Brute-force filtering densities on a spatial grid (``brute_force.py:26-136``).

    Returns ``(T, n)`` (or ``(B, T, n)`` for batched ``ys``) as a torch CUDA tensor; ``history='last'`` keeps only the
    final density ``(n,)`` / ``(B, n)``.  With ``return_nell`` also the negative log-likelihood(s) accumulated from the
    normalising constants of the updates.

    ``power_operator=True`` ('chapman' methods): the ``integration_steps`` sub-steps of a time step apply one and the same
    linear operator (``brute_force.py:115-122``), so its ``integration_steps``-th power is formed once per call (binary
    powering on the FP64 tensor-core GEMM) and every time step is one contraction instead of ``integration_steps``.  The
    same mathematics in a different order of summation (relative differences ~1e-13); off by default, where the
    reference's literal recursion runs.
    """
    import torch
    if not isinstance(drift, Drift):
        raise TypeError('drift must be a mfs_b200.functors.Drift handle (benes_drift(), well_drift(p), linear_drift(a))')
    if not isinstance(measurement_cond_pdf, MeasurementFunctor):
        raise TypeError('measurement_cond_pdf must be a MeasurementFunctor handle')
    b = dispersion.value if isinstance(dispersion, Dispersion) else float(dispersion)
    if pred_method == 'chapman-euler':
        method, order = _lib.BF_METHOD['chapman-euler'], 1
    elif pred_method.startswith('chapman-tme-'):
        method, order = _lib.BF_METHOD['chapman-tme'], int(pred_method.split('-')[-1])
    elif pred_method == 'kolmogorov':
        method, order = _lib.BF_METHOD['kolmogorov'], 1
    else:
        raise NotImplementedError(f'Prediction method {pred_method} not implemented.')   # brute_force.py:133
    if history not in ('full', 'last'):
        raise ValueError("history must be 'full' or 'last'")
    if any(np.ndim(p) > 0 for p in drift.params):
        raise ValueError('the grid operator is shared by the batch: drift parameters must be scalars')

    xs_nd = xs.dim() if isinstance(xs, torch.Tensor) else np.ndim(xs)
    if xs_nd == 2:
        # One grid PER RECORD -- the reference's own experiment: every Monte-Carlo record gets the grid spanned by its
        # moment-filter run (dardel/benes_bernoulli/brute_force.py:73-76).  The transition operator depends on the grid,
        # so there is no shared dense contraction: each record runs the reference's call shape (one record, one grid;
        # matrix-vector sub-steps on its L2-resident operator), enqueued back to back on the stream.  Results are the
        # single-record results bit for bit.
        ys_b = ys if isinstance(ys, torch.Tensor) else np.asarray(ys)
        if ys_b.ndim != 2 or ys_b.shape[0] != xs.shape[0]:
            raise ValueError('per-record grids xs (B, n) need ys (B, T) with the same B')
        ip = init_ps if isinstance(init_ps, torch.Tensor) else np.asarray(init_ps, dtype=np.float64)
        if ip.ndim != 2 or ip.shape[0] != xs.shape[0]:
            raise ValueError('per-record grids xs (B, n) need init_ps (B, n): the initial density on each grid')
        mp = measurement_cond_pdf.params
        outs, nells = [], []
        for k in range(xs.shape[0]):
            meas_k = measurement_cond_pdf.with_params(*[(np.asarray(p).reshape(-1)[k] if np.ndim(p) > 0 else p) for p in mp])
            r = brute_force_filter(drift, dispersion, meas_k, ip[k], xs[k], ys_b[k], dt, integration_steps, pred_method,
                                   history, True, device, power_operator)
            outs.append(r[0])
            nells.append(r[1])
        out, nell = torch.stack(outs), torch.stack(nells)
        return (out, nell) if return_nell else out

    if isinstance(ys, torch.Tensor) and ys.is_cuda:
        dev = ys.device
    else:
        dev = torch.device('cuda', 0 if device is None else int(device))
        ys = torch.as_tensor(np.asarray(ys)).to(dev)
    if ys.dtype == torch.bool:
        ys = ys.view(torch.uint8)
    elif ys.dtype == torch.int64:
        ys = ys.to(torch.int32)
    elif ys.dtype in (torch.float32, torch.float16):
        ys = ys.to(torch.float64)
    name = str(ys.dtype).replace('torch.', '')
    if name not in _lib.YS_DTYPE:
        raise TypeError(f'ys dtype {ys.dtype} not supported (uint8 / bool / int32 / float64)')
    batched = ys.ndim == 2
    if ys.ndim not in (1, 2):
        raise ValueError('ys must have shape (T,) or (B, T)')
    ys2 = ys.reshape(-1, ys.shape[-1]).contiguous()
    B, T = int(ys2.shape[0]), int(ys2.shape[1])

    xs_t = torch.as_tensor(np.asarray(xs.cpu() if isinstance(xs, torch.Tensor) else xs, dtype=np.float64)).to(dev)
    n = int(xs_t.shape[0])
    if xs_t.ndim != 1 or n < 3:
        raise ValueError('xs must be a 1-D grid with at least 3 points')
    ps0 = init_ps if isinstance(init_ps, torch.Tensor) else torch.as_tensor(np.asarray(init_ps, dtype=np.float64))
    ps0 = ps0.to(device=dev, dtype=torch.float64).contiguous()
    if ps0.shape == (n,):
        ps_stride = 0
    elif ps0.shape == (B, n):
        ps_stride = n
    else:
        raise ValueError(f'init_ps must have shape ({n},) or ({B}, {n}), got {tuple(ps0.shape)}')
    tprm, _ = pack_params(drift.params, ())
    mprm, mstride = pack_params(measurement_cond_pdf.params, (B,))
    tprm_t, mprm_t = torch.from_numpy(tprm).to(dev), torch.from_numpy(mprm).to(dev)

    L = _lib.lib()
    flags = _lib.BF_FLAG_POWER_OPERATOR if (power_operator and method != _lib.BF_METHOD['kolmogorov']) else 0
    ws_bytes = int(L.mfs_brute_force_workspace_bytes_ex(n, B, method, flags))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty((B, T, n) if history == 'full' else (B, n), dtype=torch.float64, device=dev)
    nell = torch.empty(B, dtype=torch.float64, device=dev) if return_nell else None

    a = _lib.BruteForceArgs()
    a.abi_version, a.pred_method, a.tme_order, a.integration_steps = _lib.ABI_VERSION, method, order, int(integration_steps)
    a.n_grid, a.drift_id, a.meas_id = n, _lib.DRIFT[drift.name], _lib.MEAS[measurement_cond_pdf.name]
    a.ys_dtype = _lib.YS_DTYPE[name]
    a.B, a.T, a.dt, a.dispersion = B, T, float(dt), b
    a.trans_params, a.meas_params, a.meas_param_stride = tprm_t.data_ptr(), mprm_t.data_ptr(), mstride
    a.xs, a.init_ps, a.init_ps_stride = xs_t.data_ptr(), ps0.data_ptr(), ps_stride
    a.ys, a.ys_stride_b, a.ys_stride_t = ys2.data_ptr(), T, 1
    a.out_mode = _lib.OUT_MODE[history]
    a.flags = flags
    a.pdfs_out = out.data_ptr()
    a.nell_out = nell.data_ptr() if nell is not None else None
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.mfs_brute_force(ctypes.byref(a), ctypes.c_void_p(stream)))
    # the workspace and staged inputs must outlive the enqueued kernels: tie them to the stream
    for t in (ws, ys2, xs_t, ps0, tprm_t, mprm_t):
        t.record_stream(torch.cuda.current_stream(dev))
    if not batched:
        out = out[0]
        nell = nell[0] if nell is not None else None
    return (out, nell) if return_nell else out

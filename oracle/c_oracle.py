"""ctypes loader of ``oracle/libmfs_oracle.so`` (C restatement of the reference algorithm).  TEST INFRASTRUCTURE ONLY:
importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs, never from ``mfs_b200``.

``filter_1d`` takes the same functor handles / arrays as ``mfs_b200.one_dim.filtering`` and fills the product's own
``mfs_filter1d_args`` struct with HOST pointers, so the CUDA path and the oracle see identical arguments.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, 'libmfs_oracle.so')
_lib = None


def build():
    proc = subprocess.run(['make', '-C', _HERE], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('building the C oracle failed:\n' + proc.stdout + proc.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.mfs_oracle_filter_1d.restype = ctypes.c_int
        _lib.mfs_oracle_moment_quadrature_1d.restype = ctypes.c_int
        _lib.mfs_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().mfs_oracle_max_threads())


def moment_quadrature(ms, mean=None, scale=None, ldl=False):
    ms = np.ascontiguousarray(ms, dtype=np.float64)
    n = ms.shape[-1] // 2
    B = int(np.prod(ms.shape[:-1])) if ms.ndim > 1 else 1
    w, x = np.empty((B, n)), np.empty((B, n))
    as_p = lambda a: None if a is None else np.ascontiguousarray(
        np.broadcast_to(np.asarray(a, dtype=np.float64), ms.shape[:-1]).reshape(B)).ctypes.data_as(ctypes.c_void_p)
    mean_c = None if mean is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mean, float), ms.shape[:-1]).reshape(B))
    scale_c = None if scale is None else np.ascontiguousarray(np.broadcast_to(np.asarray(scale, float), ms.shape[:-1]).reshape(B))
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    lib().mfs_oracle_moment_quadrature_1d(ctypes.c_int(n), ctypes.c_int64(B), p(ms), p(mean_c), p(scale_c),
                                          ctypes.c_int(int(ldl)), p(w), p(x))
    return w.reshape(ms.shape[:-1] + (n,)), x.reshape(ms.shape[:-1] + (n,))


def filter_1d(mode, transition, measurement, ms0, ys, mean0=None, scale0=None, stable=False, history='full',
              num_threads=0):
    """Run the C oracle.  ``transition`` is any member of a ``mfs_b200`` factory tuple (only its spec is used),
    ``measurement`` a ``MeasurementFunctor``.  Returns a dict like ``mfs_b200.one_dim.filtering._run``."""
    from mfs_b200 import _lib as P  # struct mirror only
    from mfs_b200.functors import pack_params
    spec = transition.spec
    ys = np.asarray(ys)
    if ys.dtype == np.bool_:
        ys = ys.view(np.uint8)
    elif ys.dtype == np.int64:
        ys = ys.astype(np.int32)
    batch_shape = tuple(ys.shape[:-1])
    T = ys.shape[-1]
    B = int(np.prod(batch_shape)) if batch_shape else 1
    ys_c = np.ascontiguousarray(ys.reshape(B, T))
    ms0 = np.asarray(ms0, dtype=np.float64)
    M = ms0.shape[-1] // 2 * 2
    N = M // 2

    def per_filter(arr, trailing):
        arr = np.asarray(arr, dtype=np.float64)
        if arr.shape == trailing:
            return np.ascontiguousarray(arr.reshape((1,) + trailing)), 0
        return np.ascontiguousarray(np.broadcast_to(arr, batch_shape + trailing).reshape((B,) + trailing)), \
            int(np.prod(trailing)) if trailing else 1

    ptr = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    a = P.Filter1dArgs()
    a.abi_version, a.mode, a.N, a.stable = P.ABI_VERSION, P.MODE[mode], N, int(bool(stable))
    a.B, a.T = B, T
    a.trans_id, a.drift_id = P.TRANS[spec.family], P.DRIFT[spec.drift.name]
    a.tme_order, a.meas_id = spec.order, P.MEAS[measurement.name]
    a.dt, a.dispersion = spec.dt, spec.dispersion
    tprm, a.trans_param_stride = pack_params(spec.packed_params(), batch_shape)
    mprm, a.meas_param_stride = pack_params(measurement.params, batch_shape)
    ms0_tab, a.ms0_stride = per_filter(ms0[..., :M], (M,))
    a.trans_params, a.meas_params, a.ms0 = ptr(tprm), ptr(mprm), ptr(ms0_tab)
    keep = [tprm, mprm, ms0_tab, ys_c]
    if mode != 'raw':
        m0, a.mean0_stride = per_filter(mean0, ())
        a.mean0 = ptr(m0)
        keep.append(m0)
    if mode == 'scaled':
        s0, a.scale0_stride = per_filter(scale0, ())
        a.scale0 = ptr(s0)
        keep.append(s0)
    a.ys, a.ys_dtype = ptr(ys_c), P.YS_DTYPE[str(ys_c.dtype)]
    a.ys_stride_b, a.ys_stride_t = T, 1
    a.out_mode = P.OUT_MODE[history]
    nell, status = np.empty(B), np.empty(B, dtype=np.int32)
    ms_out = mean_out = scale_out = None
    if history == 'full':
        ms_out = np.empty((B, T, M))
        a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = T * M, M, T
        aux_shape = (B, T)
    elif history == 'last':
        ms_out = np.empty((B, M))
        a.ms_stride_b, a.ms_stride_t, a.aux_stride_b = M, 0, 1
        aux_shape = (B,)
    if ms_out is not None:
        a.ms_out = ptr(ms_out)
        if mode != 'raw':
            mean_out = np.empty(aux_shape)
            a.mean_out = ptr(mean_out)
        if mode == 'scaled':
            scale_out = np.empty(aux_shape)
            a.scale_out = ptr(scale_out)
    a.nell_out, a.status_out = ptr(nell), ptr(status)
    used = lib().mfs_oracle_filter_1d(ctypes.byref(a), ctypes.c_int(num_threads))
    if used < 0:
        raise RuntimeError('mfs_oracle_filter_1d rejected its arguments')
    tail = (T,) if history == 'full' else ()
    rs = lambda t, tl: None if t is None else t.reshape(batch_shape + tl)
    return {'ms': rs(ms_out, tail + (M,)), 'mean': rs(mean_out, tail), 'scale': rs(scale_out, tail),
            'nell': rs(nell, ()), 'status': rs(status, ()), 'threads': used}

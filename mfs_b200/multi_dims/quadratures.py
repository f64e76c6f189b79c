"""Host-side mirror of ``mfs/multi_dims/quadratures.py``: batched ``moment_quadrature_nd`` on the GPU (d = 2)."""
import ctypes

import numpy as np

from .. import _lib

__all__ = ['moment_quadrature_nd', 'nd_cartesian_prod_indices']


def nd_cartesian_prod_indices(d: int, n: int) -> np.ndarray:
    """All index tuples of ``{0..n-1}^d`` in the reference's order (``quadratures.py:29-48``)."""
    grids = np.meshgrid(*([np.arange(n)] * d), indexing='ij')
    return np.stack([g.reshape(-1) for g in grids], axis=1).astype(np.int64)


def moment_quadrature_nd(ms, inds, mean=None, scale=None, ldl: bool = False):
    """Mirror of ``mfs/multi_dims/quadratures.py:120-178``: ``ms`` ``(z,)`` or ``(..., z)`` moments in graded-lex order,
    ``inds`` the ``(3, s, s)`` table of ``gram_and_hankel_indices_graded_lexico(N, 2)``; returns ``weights (..., s**2)``,
    ``nodes (..., s**2, 2)`` (``nodes * scale + mean`` when given).  The order of the eigenpairs inside each of the two
    eigen-decompositions (hence of the nodes) and the eigenvector signs are arbitrary, as with any ``eigh``.  torch CUDA
    in -> torch CUDA out, NumPy in -> NumPy out; a non-positive-definite Gram matrix gives NaN (``ldl=False``)."""
    import torch
    inds = np.asarray(inds)
    if inds.ndim != 3 or inds.shape[0] != 3:
        raise _lib.MfsError(f'only d = 2 is implemented (inds has shape {inds.shape})')
    s = inds.shape[1]
    N = int(round((np.sqrt(8 * s + 1) - 1) / 2))
    z = N * (2 * N + 1)
    is_np = not (isinstance(ms, torch.Tensor) and ms.is_cuda)
    dev = torch.device('cuda', torch.cuda.current_device()) if is_np else ms.device
    ms_t = torch.as_tensor(np.asarray(ms, dtype=np.float64) if is_np else ms, dtype=torch.float64, device=dev)
    if ms_t.shape[-1] != z:
        raise ValueError(f'expected {z} moments for N = {N}, got {ms_t.shape[-1]}')
    batch_shape = tuple(ms_t.shape[:-1])
    B = int(np.prod(batch_shape)) if batch_shape else 1
    ms_c = ms_t.reshape(B, z).contiguous()

    def aux(v):
        if v is None:
            return None
        t = torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                            dtype=torch.float64, device=dev)
        return t.expand(batch_shape + (2,)).reshape(B, 2).contiguous()

    mean_t, scale_t = aux(mean), aux(scale)
    inds_t = torch.from_numpy(np.ascontiguousarray(inds.astype(np.int32))).to(dev)
    w = torch.empty((B, s * s), dtype=torch.float64, device=dev)
    x = torch.empty((B, s * s, 2), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().mfs_moment_quadrature_nd(
            N, 2, B, ms_c.data_ptr(), None if mean_t is None else mean_t.data_ptr(),
            None if scale_t is None else scale_t.data_ptr(), inds_t.data_ptr(), int(bool(ldl)),
            w.data_ptr(), x.data_ptr(), ctypes.c_void_p(stream)))
    w, x = w.reshape(batch_shape + (s * s,)), x.reshape(batch_shape + (s * s, 2))
    if is_np:
        return w.cpu().numpy(), x.cpu().numpy()
    return w, x

"""Host-side mirror of ``mfs/one_dim/quadtures.py``: batched ``moment_quadrature`` on the GPU."""
import ctypes
import math

import numpy as np

from .. import _lib

__all__ = ['moment_quadrature', 'hankel_indices']


def hankel_indices(n: int):
    """Index tables of the Hankel pair (``mfs/one_dim/quadtures.py:29-60``); kept for API parity (host, NumPy)."""
    inds = np.arange(n)[:, None] + np.arange(n)[None, :]
    return inds, inds + 1


def moment_quadrature(ms, mean=0., scale=1., sort_nodes: bool = False, ldl: bool = False):
    """Gauss quadrature (weights, nodes) from moments, mirror of ``mfs/one_dim/quadtures.py:83-133``.

    ``ms`` is ``(..., 2n)`` as a torch CUDA tensor (device in, device out) or a NumPy array (copied to the current
    CUDA device and back).  ``mean`` / ``scale`` are scalars or ``(...)`` arrays.
    """
    import torch
    is_np = not (isinstance(ms, torch.Tensor) and ms.is_cuda)
    dev = torch.device('cuda', torch.cuda.current_device()) if is_np else ms.device
    ms_t = torch.as_tensor(np.asarray(ms, dtype=np.float64) if is_np else ms, dtype=torch.float64, device=dev)
    n = math.floor(ms_t.shape[-1] / 2)
    batch_shape = tuple(ms_t.shape[:-1])
    B = int(np.prod(batch_shape)) if batch_shape else 1
    ms_c = ms_t[..., :2 * n].reshape(B, 2 * n).contiguous()

    def aux(v, default):
        if np.ndim(v) == 0 and float(v) == default:
            return None
        t = torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                            dtype=torch.float64, device=dev)
        return t.expand(batch_shape).reshape(B).contiguous()

    mean_t, scale_t = aux(mean, 0.), aux(scale, 1.)
    w = torch.empty((B, n), dtype=torch.float64, device=dev)
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().mfs_moment_quadrature_1d(
            n, B, ms_c.data_ptr(), None if mean_t is None else mean_t.data_ptr(),
            None if scale_t is None else scale_t.data_ptr(), int(bool(sort_nodes)), int(bool(ldl)),
            w.data_ptr(), x.data_ptr(), ctypes.c_void_p(stream)))
    w, x = w.reshape(batch_shape + (n,)), x.reshape(batch_shape + (n,))
    if is_np:
        return w.cpu().numpy(), x.cpu().numpy()
    return w, x

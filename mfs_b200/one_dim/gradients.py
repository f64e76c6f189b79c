"""nell and its gradient w.r.t. model parameters: the objective the reference hands to L-BFGS-B.

The reference differentiates ``nell`` through the filter scan with ``jax.grad`` (``dardel/parameter_estimation/
mf.py:37-73``: ``obj_func`` -> ``jaxopt.ScipyMinimize(method='L-BFGS-B')``; "differentiable in the parameter",
``README.md:45``).  Here the derivative is carried forward through the scan by the kernel of
``mfs_b200/csrc/filter1d_grad.cuh`` (dual numbers: Hankel recurrence, QL eigen-solve, transition moments and
likelihood), for a whole batch of independent filters -- theta grids, Monte-Carlo runs, or all line-search points of a
batched optimiser -- in one launch.

``wrt`` names the parameters: ``('drift', k)`` is the k-th packed parameter of the transition functor (well: theta1;
linear: a; normal_affine: F, Sigma), ``('meas', k)`` the k-th parameter of the measurement functor (poisson_softplus:
theta2; bernoulli_logistic_cubic: c0, c1; gaussian: h, r).  Default: every parameter the two handles carry.
"""
from typing import Optional, Sequence

from .. import _lib
from .filtering import _run, _check_transition, _check_measurement

__all__ = ['moment_filter_rms_value_and_grad', 'moment_filter_cms_value_and_grad']


def _tangent_ids(spec, meas, wrt):
    if wrt is None:
        wrt = [('drift', k) for k in range(len(spec.packed_params()))] + [('meas', k) for k in range(len(meas.params))]
    ids = []
    for kind, k in wrt:
        n = len(spec.packed_params()) if kind == 'drift' else len(meas.params) if kind == 'meas' else -1
        if n < 0:
            raise ValueError(f"wrt entries are ('drift', k) or ('meas', k), got {(kind, k)!r}")
        if not 0 <= int(k) < n:
            raise ValueError(f'{kind} functor has {n} parameter(s); index {k} is out of range')
        ids.append(int(k) + (_lib.MAX_PARAMS if kind == 'meas' else 0))
    if not ids:
        raise ValueError('no parameter to differentiate with respect to')
    return ids


def moment_filter_rms_value_and_grad(state_cond_raw_moments, measurement_cond_pdf, rms0, ys,
                                     wrt: Optional[Sequence] = None, *, return_status: bool = False):
    """``jax.value_and_grad`` of ``lambda theta: moment_filter_rms(...)[1]`` (``mfs/one_dim/filtering.py:32-89``).
    ``ys``: CUDA tensor ``(..., T)``.  Returns ``(nell (...), grad (..., len(wrt)))`` (+ status)."""
    fn = _check_transition(state_cond_raw_moments, 'raw', 'state_cond_raw_moments')
    meas = _check_measurement(measurement_cond_pdf)
    out = _run('raw', fn.spec, meas, rms0, None, None, ys, False, 'none', None, return_status, 0,
               grad_ids=_tangent_ids(fn.spec, meas, wrt))
    res = (out['nell'], out['grad'])
    return res + (out['status'],) if return_status else res


def moment_filter_cms_value_and_grad(state_cond_central_moments, state_cond_mean, measurement_cond_pdf, cms0, mean0,
                                     ys, wrt: Optional[Sequence] = None, *, return_status: bool = False):
    """``jax.value_and_grad`` of the central-moment objective (``dardel/parameter_estimation/mf.py:52-54``)."""
    fn = _check_transition(state_cond_central_moments, 'central', 'state_cond_central_moments')
    fm = _check_transition(state_cond_mean, 'mean', 'state_cond_mean')
    if fm.spec is not fn.spec:
        raise ValueError('state_cond_central_moments and state_cond_mean must come from the same factory call')
    meas = _check_measurement(measurement_cond_pdf)
    out = _run('central', fn.spec, meas, cms0, mean0, None, ys, False, 'none', None, return_status, 0,
               grad_ids=_tangent_ids(fn.spec, meas, wrt))
    res = (out['nell'], out['grad'])
    return res + (out['status'],) if return_status else res

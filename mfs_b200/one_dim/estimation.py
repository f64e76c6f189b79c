"""Batched maximum-likelihood parameter estimation on top of the nell + gradient kernel.

The reference estimates (theta1, theta2) of the well--Poisson model one Monte-Carlo run per process:
``jaxopt.ScipyMinimize(method='L-BFGS-B', fun=obj_func)`` with ``obj_func(params, ys) = nell`` of
``moment_filter_cms`` at ``softplus(params)`` (``dardel/parameter_estimation/mf.py:37-73``,
``run_parameter_estimation_mf.sh``: 1000 runs).  On one GPU all runs advance in lockstep instead: every objective /
gradient evaluation of every run is ONE launch of ``mfs_filter_1d_grad`` over the batch, and the quasi-Newton update
(dense BFGS -- the problems have 1-4 parameters -- with Armijo backtracking) is a handful of batched tensor
operations.  ``batched_bfgs`` is generic (any ``fun(x (B, P)) -> (f (B,), g (B, P))`` on any torch device);
``estimate_well_poisson`` wires it to the reference's objective.
"""
from typing import Callable, NamedTuple

import torch

__all__ = ['batched_bfgs', 'BfgsResult', 'softplus_objective', 'estimate_well_poisson']


class BfgsResult(NamedTuple):
    x: torch.Tensor            # (B, P) minimisers
    fun: torch.Tensor          # (B,) objective values
    grad: torch.Tensor         # (B, P)
    success: torch.Tensor      # (B,) bool: converged by gtol / ftol
    nit: torch.Tensor          # (B,) iterations taken
    nfev: int                  # batched objective evaluations (= kernel launches of the objective)


def batched_bfgs(fun: Callable, x0: torch.Tensor, maxiter: int = 100, gtol: float = 1e-5, ftol: float = 2.2e-9,
                 max_linesearch: int = 25, c1: float = 1e-4, c2: float = 0.9) -> BfgsResult:
    """Minimise B independent smooth functions of P variables in lockstep.

    ``fun(x)`` evaluates all B problems at once; a non-finite value (e.g. a filter that lost positive definiteness at
    a trial point) is treated as "step too long".  Stopping rules are those of SciPy's L-BFGS-B defaults
    (``max|g| <= gtol`` or relative decrease ``<= ftol``).
    """
    x = x0.clone()
    B, P = x.shape
    f, g = fun(x)
    nfev = 1
    eye = torch.eye(P, dtype=x.dtype, device=x.device).expand(B, P, P)
    H = eye.clone()
    done = ~torch.isfinite(f) | (g.abs().amax(dim=1) <= gtol)
    success = torch.isfinite(f) & (g.abs().amax(dim=1) <= gtol)
    nit = torch.zeros(B, dtype=torch.int64, device=x.device)
    first = True
    for _ in range(maxiter):
        if bool(done.all()):
            break
        d = -torch.einsum('bij,bj->bi', H, g)
        slope = (g * d).sum(dim=1)
        reset = ~(slope < 0)                      # not a descent direction (or NaN): restart from steepest descent
        if bool(reset.any()):
            H = torch.where(reset[:, None, None], eye, H)
            d = torch.where(reset[:, None], -g, d)
            slope = torch.where(reset, -(g * g).sum(dim=1), slope)
        t = torch.ones(B, dtype=x.dtype, device=x.device)
        if first:                                  # first step: unit length at most
            t = torch.minimum(t, 1.0 / g.norm(dim=1).clamp_min(1e-300))
            first = False
        # Lockstep line search for the (weak) Wolfe conditions: halve while the Armijo test fails (or the trial value is
        # not finite), DOUBLE while it holds but the slope at the trial point is still steeper than c2 * slope -- the
        # extrapolation SciPy's dcsrch performs, without which a start in a flat region settles in the nearest dip.
        accepted = done.clone()
        x_new, f_new, g_new = x.clone(), f.clone(), g.clone()
        have_cand = torch.zeros_like(done)
        shrunk = torch.zeros_like(done)
        for _ls in range(max_linesearch):
            trial = torch.where(accepted[:, None], x_new, x + t[:, None] * d)
            ft, gt = fun(trial)
            nfev += 1
            finite = torch.isfinite(ft) & torch.isfinite(gt).all(dim=1)
            armijo = ~accepted & finite & (ft <= f + c1 * t * slope)
            armijo = armijo & (~have_cand | (ft < f_new))            # an expansion must keep improving on its candidate
            curv = (gt * d).sum(dim=1) >= c2 * slope
            take = armijo                                             # new best point along the ray
            x_new = torch.where(take[:, None], trial, x_new)
            f_new = torch.where(take, ft, f_new)
            g_new = torch.where(take[:, None], gt, g_new)
            fail = ~accepted & ~armijo
            # finished: Wolfe point reached, an expansion stopped paying, or Armijo holds after a contraction
            finish = (armijo & (curv | shrunk)) | (fail & have_cand)
            have_cand = have_cand | armijo
            accepted = accepted | finish
            if bool(accepted.all()):
                break
            expand = armijo & ~finish
            t = torch.where(expand, 2.0 * t, torch.where(fail & ~accepted, 0.5 * t, t))
            shrunk = shrunk | (fail & ~accepted)
        accepted = accepted | have_cand
        stalled = ~accepted                       # line search failed: stop this problem where it is
        s, yv = x_new - x, g_new - g
        sy = (s * yv).sum(dim=1)
        upd = (sy > 1e-12 * s.norm(dim=1) * yv.norm(dim=1)) & ~done & accepted
        rho = torch.where(upd, 1.0 / torch.where(upd, sy, torch.ones_like(sy)), torch.zeros_like(sy))
        V = eye - rho[:, None, None] * s[:, :, None] * yv[:, None, :]
        H_new = V @ H @ V.transpose(1, 2) + rho[:, None, None] * s[:, :, None] * s[:, None, :]
        H = torch.where(upd[:, None, None], H_new, H)
        small_g = g_new.abs().amax(dim=1) <= gtol
        small_f = (f - f_new) <= ftol * torch.maximum(torch.maximum(f.abs(), f_new.abs()), torch.ones_like(f))
        moved = ~done & accepted
        nit = nit + moved.to(nit.dtype)
        success = success | (moved & (small_g | small_f))
        x, f, g = x_new, f_new, g_new
        done = done | stalled | small_g | (moved & small_f)
    return BfgsResult(x, f, g, success, nit, nfev)


def softplus_objective(value_and_grad: Callable) -> Callable:
    """The reference optimises unconstrained ``params`` with ``theta = log(exp(params) + 1)``
    (``dardel/parameter_estimation/mf.py:39``); chain rule: d/dparams = d/dtheta * sigmoid(params)."""
    def fun(raw):
        theta = torch.nn.functional.softplus(raw)
        f, g = value_and_grad(theta)
        return f, g * torch.sigmoid(raw)
    return fun


def estimate_well_poisson(ys, N: int, init=(0.1, 0.1), euler: bool = False, maxiter: int = 100):
    """``dardel/parameter_estimation/mf.py`` for a batch of measurement records ``ys`` (CUDA int32 tensor ``(B, T)``):
    central-moment filter, TME-normal order 2 (or Euler--Maruyama with ``euler=True``), start at ``init`` (the
    reference's ``--p1_init/--p2_init`` = 0.1).  Returns ``(theta_hat (B, 2), BfgsResult)``."""
    from .gradients import moment_filter_cms_value_and_grad
    from .moments import sde_cond_moments_tme_normal, sde_cond_moments_euler
    from .ss_models import well_poisson
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., N)

    def value_and_grad(theta):
        th = theta.detach().cpu().numpy()
        fam = sde_cond_moments_euler(drift(th[:, 0]), disp, dt, N) if euler else \
            sde_cond_moments_tme_normal(drift(th[:, 0]), disp, dt, 2, N)
        return moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(th[:, 1]), ic.cms, ic.mean, ys)

    B = ys.shape[0]
    init_t = torch.tensor(init, dtype=torch.float64, device=ys.device)
    raw0 = torch.log(torch.expm1(init_t)).expand(B, 2).contiguous()       # mf.py:68
    res = batched_bfgs(softplus_objective(value_and_grad), raw0, maxiter=maxiter)
    return torch.nn.functional.softplus(res.x), res

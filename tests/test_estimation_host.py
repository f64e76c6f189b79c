"""Host logic of the batched optimiser (CPU tensors, analytic objectives): the lockstep BFGS reaches the minimisers
SciPy's L-BFGS-B finds, handles non-finite trial values, and the softplus chain rule is the reference's
(dardel/parameter_estimation/mf.py:39, 68-71)."""
import numpy as np
import torch
from scipy.optimize import minimize

from mfs_b200.one_dim.estimation import batched_bfgs, softplus_objective


def _rosenbrock(a, b):
    def fun(x):
        f = (a - x[:, 0]) ** 2 + b * (x[:, 1] - x[:, 0] ** 2) ** 2
        g = torch.stack([-2 * (a - x[:, 0]) - 4 * b * x[:, 0] * (x[:, 1] - x[:, 0] ** 2), 2 * b * (x[:, 1] - x[:, 0] ** 2)], 1)
        return f, g
    return fun


def test_batched_bfgs_matches_scipy_on_rosenbrock_family():
    a = torch.tensor([1., 0.5, 1.5, -1.], dtype=torch.float64)
    b = torch.tensor([10., 5., 20., 3.], dtype=torch.float64)
    x0 = torch.tensor([[-1.2, 1.], [0., 0.], [2., 2.], [0.3, -0.7]], dtype=torch.float64)
    res = batched_bfgs(_rosenbrock(a, b), x0, maxiter=200, gtol=1e-8, ftol=0.)
    assert bool(res.success.all())
    for k in range(4):
        ref = minimize(lambda v: float((a[k] - v[0]) ** 2 + b[k] * (v[1] - v[0] ** 2) ** 2), x0[k].numpy(), method='L-BFGS-B',
                       options=dict(ftol=1e-15, gtol=1e-10))
        np.testing.assert_allclose(res.x[k].numpy(), ref.x, atol=1e-5)
        np.testing.assert_allclose(res.x[k].numpy(), [float(a[k]), float(a[k]) ** 2], atol=1e-6)
    assert res.nfev < 400 and int(res.nit.max()) <= 200


def test_non_finite_trial_points_shrink_the_step():
    """f = x^2 - log(3 - x) on x < 3 and NaN beyond (a filter that diverged at a trial theta)."""
    def fun(x):
        inside = x[:, 0] < 3
        f = torch.where(inside, x[:, 0] ** 2 - torch.log((3 - x[:, 0]).clamp_min(1e-300)), torch.full_like(x[:, 0], float('nan')))
        g = torch.where(inside, 2 * x[:, 0] + 1 / (3 - x[:, 0]), torch.full_like(x[:, 0], float('nan')))[:, None]
        return f, g
    res = batched_bfgs(fun, torch.tensor([[2.9], [-40.], [2.999]], dtype=torch.float64), gtol=1e-9, ftol=0.)
    xstar = (3 - np.sqrt(11)) / 2          # 2x + 1/(3-x) = 0
    assert bool(res.success.all())
    np.testing.assert_allclose(res.x[:, 0].numpy(), xstar, atol=1e-7)
    # a start outside the domain is reported as a failure, not a crash
    bad = batched_bfgs(fun, torch.tensor([[5.]], dtype=torch.float64))
    assert not bool(bad.success[0])


def test_softplus_chain_rule():
    def vg(theta):
        return (theta ** 2).sum(1), 2 * theta
    raw = torch.tensor([[0.3, -1.2], [2., 0.1]], dtype=torch.float64, requires_grad=True)
    f, g = softplus_objective(vg)(raw.detach())
    ref = (torch.nn.functional.softplus(raw) ** 2).sum()
    ref.backward()
    np.testing.assert_allclose(g.numpy(), raw.grad.numpy(), rtol=1e-12)
    # the reference's initialisation maps back to its initial guess (mf.py:68)
    init = torch.tensor([0.1, 0.1], dtype=torch.float64)
    np.testing.assert_allclose(torch.nn.functional.softplus(torch.log(torch.expm1(init))).numpy(), 0.1, rtol=1e-12)

"""The reference's parameter-estimation experiment (dardel/parameter_estimation/mf.py, run_parameter_estimation_mf.sh:
N = 7, 1000 Monte-Carlo runs, one OS process each) as ONE batched job on a GPU: simulate the records at theta = (3, 3),
run all maximum-likelihood fits in lockstep (batched BFGS on the nell + gradient kernel), report theta-hat.

    python tools/parameter_estimation.py [n_runs] [N] [T]
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy.optimize import minimize
from mfs_b200.simulate import simulate_1d
from mfs_b200.one_dim.estimation import estimate_well_poisson
from mfs_b200.one_dim.gradients import moment_filter_cms_value_and_grad
from mfs_b200.one_dim.moments import sde_cond_moments_tme_normal
from mfs_b200.one_dim.ss_models import well_poisson

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 7
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dt, _, _, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
_, _, ys = simulate_1d(drift(3.), disp, dt, T, ic, pmf(3.), runs, 2024)     # mf.py:58-65 (100 TME-3 sub-steps)
torch.cuda.synchronize()
estimate_well_poisson(ys[:8], N, maxiter=2)                                  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
theta, res = estimate_well_poisson(ys, N)
torch.cuda.synchronize()
el = time.perf_counter() - t0
th = theta.cpu().numpy()
ok = res.success.cpu().numpy()
print(f'# Parameter estimation, well-Poisson, central moments, TME-normal order 2, N = {N}, T = {T}, {runs} Monte-Carlo runs at theta = (3, 3)')
print(f'batched BFGS: {el:.2f} s wall for all runs, {res.nfev} batched objective+gradient evaluations '
      f'({res.nfev * runs * T / el:.3e} filter-steps/s incl. host logic), iterations median {int(res.nit.median())} max {int(res.nit.max())}, '
      f'converged {ok.mean():.3f}')
print(f'theta-hat: median ({np.median(th[ok, 0]):.3f}, {np.median(th[ok, 1]):.3f}), mean ({th[ok, 0].mean():.3f}, {th[ok, 1].mean():.3f}), '
      f'mean abs error ({np.abs(th[ok, 0] - 3).mean():.3f}, {np.abs(th[ok, 1] - 3).mean():.3f}), '
      f'max |grad| at the solutions median {float(res.grad.abs().amax(1).median()):.1e}')

# cross-check: SciPy's L-BFGS-B (what jaxopt.ScipyMinimize wraps, mf.py:69-70) on the same objective, a few runs
def objective(raw, k):
    th_ = np.log1p(np.exp(raw))
    fam = sde_cond_moments_tme_normal(drift(th_[0]), disp, dt, 2, N)
    f, g = moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(th_[1]), ic.cms, ic.mean, ys[k:k + 1])
    f, g = float(f[0]), g[0].cpu().numpy() / (1 + np.exp(-raw))
    return (f, g) if np.isfinite(f) else (1e300, np.zeros(2))

raw0 = np.log(np.expm1(np.array([0.1, 0.1])))
worst = 0.
t0 = time.perf_counter()
for k in range(min(5, runs)):
    r = minimize(objective, raw0, args=(k,), jac=True, method='L-BFGS-B')
    worst = max(worst, float(np.max(np.abs(np.log1p(np.exp(r.x)) - th[k]))))
    print(f'run {k}: scipy L-BFGS-B theta-hat {np.log1p(np.exp(r.x))} nell {r.fun:.6f} ({r.nfev} evaluations) | batched {th[k]} nell {float(res.fun[k]):.6f}')
print(f'max |theta-hat difference| vs SciPy on these runs: {worst:.2e}; SciPy took {(time.perf_counter() - t0) / min(5, runs):.2f} s per run')

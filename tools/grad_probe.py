"""Throughput of the nell + gradient kernel next to the value kernel (CUDA events, second of two runs):
well--Poisson, central, TME-normal order 2, N = 7 (dardel/run_parameter_estimation_mf.sh:34), T = 1000."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mfs_b200.simulate import simulate_1d
from mfs_b200.one_dim.filtering import moment_filter_cms
from mfs_b200.one_dim.gradients import moment_filter_cms_value_and_grad
from mfs_b200.one_dim.moments import sde_cond_moments_tme_normal
from mfs_b200.one_dim.ss_models import well_poisson

N = int(sys.argv[1]) if len(sys.argv) > 1 else 7
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 64 * 8
T = 1000
dt, _, _, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
_, _, ys = simulate_1d(drift(3.), disp, dt, T, ic, pmf(3.), B, 7)
fam = sde_cond_moments_tme_normal(drift(3.), disp, dt, 2, N)


def timed(fn):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), r


ms_v, rv = timed(lambda: moment_filter_cms(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, ys, history='none', return_status=True))
ms_g, rg = timed(lambda: moment_filter_cms_value_and_grad(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, ys, return_status=True))
ok = (rv[3] < 0) & (rg[2] < 0)
print(f'N={N} B={B} T={T} well-poisson central tme_normal-2')
print(f'value kernel        : {ms_v:9.3f} ms  {B * T / ms_v * 1e3:.3e} filter-steps/s  diverged {float((rv[3] >= 0).double().mean()):.4f}')
print(f'value + 2 gradients : {ms_g:9.3f} ms  {B * T / ms_g * 1e3:.3e} filter-steps/s  diverged {float((rg[2] >= 0).double().mean()):.4f}  '
      f'(x{ms_g / ms_v:.1f} the value kernel)')
print(f'max rel nell difference on common survivors: {float(((rv[2][ok] - rg[0][ok]).abs() / rv[2][ok].abs()).max()):.2e}')
print(f'mean gradient over trajectories at theta=(3,3): {rg[1][ok].mean(0).cpu().numpy()} (per-trajectory std {rg[1][ok].std(0).cpu().numpy()})')

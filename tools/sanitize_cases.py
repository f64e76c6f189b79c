"""Small invocations of every kernel of the path for `compute-sanitizer` (memcheck / racecheck / initcheck are 10-100x
slower than a plain run, so the batches are tiny): the 1-D filter in all modes with segmented execution, NaN tails,
meanvar history, time-chunked carry and T = 0; the batched quadrature and characteristic function; the 2-D filter and
quadrature; the grid filter on its GEMV and DMMA-GEMM paths; the simulators; the nell + gradient kernel.
usage: compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms, moment_filter_scms
from mfs_b200.one_dim.moments import sde_cond_moments_tme, sde_cond_moments_tme_normal, sde_cond_moments_euler, characteristic_fn
from mfs_b200.one_dim.quadtures import moment_quadrature
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson
from mfs_b200.one_dim.gradients import moment_filter_cms_value_and_grad
from mfs_b200.simulate import simulate_1d, simulate_prey_predator
from mfs_b200 import synthetic

N = 8
dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
fam = sde_cond_moments_tme(drift, disp, dt, 3)
ys = synthetic.benes_bernoulli_ys_torch(300, 300, 5, 'cuda')        # T = 300: 5 segments of 64 steps, some filters diverge
for hist in ('full', 'meanvar', 'last', 'none'):
    out = moment_filter_rms(fam[0], pmf, ic.rms, ys, history=hist, return_status=True)
print('1-D raw ok, diverged', int((out[-1] >= 0).sum()))
moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys[:, :140].contiguous(), return_status=True)
moment_filter_scms(fam[2], fam[4], pmf, ic.scms, ic.mean, np.sqrt(ic.variance), ys[:, :70].contiguous(), stable=True)
a = moment_filter_rms(fam[0], pmf, ic.rms, ys[:, :150].contiguous(), return_carry=True, recompute_predict_quadrature=True)
moment_filter_rms(fam[0], pmf, ic.rms, ys[:, 150:].contiguous(), carry=a[-1], t_offset=150, recompute_predict_quadrature=True)
moment_filter_rms(fam[0], pmf, ic.rms, ys[:, :0], history='last')
moment_filter_rms(fam[0], pmf, ic.rms, ys.cpu().numpy()[:, :130], chunk_filters=128, device=0)      # host pipeline
e = sde_cond_moments_euler(drift, disp, dt, 5)
dt5, _, _, ic5, _, _, _, _, _ = benes_bernoulli(5)
moment_filter_scms(e[2], e[4], pmf, ic5.scms, ic5.mean, np.sqrt(ic5.variance), ys[:64, :50].contiguous())
print('1-D modes ok')
dtw, _, _, icw, driftw, dispw, _, pmfw, _ = well_poisson(3., 7)
th1, th2 = np.linspace(2., 4., 5), np.linspace(2., 4., 5)
ysw = simulate_1d(driftw(3.), dispw, dtw, 200, icw, pmfw(3.), 40, 670)[2]
famw = sde_cond_moments_tme_normal(driftw(th1[:, None]), dispw, dtw, 2, 7)
moment_filter_cms(famw[1], famw[3], pmfw(th2[:, None]), icw.cms, icw.mean, ysw, history='none')   # theta grid over shared records
fam1 = sde_cond_moments_tme_normal(driftw(3.), dispw, dtw, 2, 7)
moment_filter_cms_value_and_grad(fam1[1], fam1[3], pmfw(3.), icw.cms, icw.mean, ysw[:, :60].contiguous())
print('theta grid + gradient ok')
w, x = moment_quadrature(torch.from_numpy(np.tile(ic.rms[:16], (70, 1))).cuda(), sort_nodes=True)
characteristic_fn(np.linspace(-3., 3., 50), ic.rms[:16])
print('quadrature ok')

from mfs_b200.multi_dims.multi_indices import generate_graded_lexico_multi_indices, gram_and_hankel_indices_graded_lexico
from mfs_b200.multi_dims.filtering import moment_filter_nd_cms, moment_filter_nd_rms
from mfs_b200.multi_dims.moments import sde_cond_moments_tme_normal as tn_nd, sde_cond_moments_tme as t_nd
from mfs_b200.multi_dims.quadratures import moment_quadrature_nd
from mfs_b200.multi_dims.ss_models import prey_predator
for Nn, B2, T2 in ((3, 9, 6), (5, 6, 4), (7, 5, 2)):
    mis = generate_graded_lexico_multi_indices(2, 2 * Nn - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(Nn, 2)
    dt2, _, _, gs, drift2, disp2, _, pmf2, _ = prey_predator(mis)
    ys2 = simulate_prey_predator(drift2, disp2, dt2, T2, gs, pmf2, B2, 677, integration_steps=5)[2]
    f = tn_nd(drift2, disp2, dt2, 2, mis)
    moment_filter_nd_cms((f[1], 'index'), f[3], pmf2, ys2, (mis, inds), gs.cms, gs.mean, return_status=True)
    f = t_nd(drift2, disp2, dt2, 2)
    moment_filter_nd_rms((f[0], 'multi-index'), pmf2, ys2, (mis, inds), gs.rms, history='last')
    moment_quadrature_nd(gs.cms, inds, mean=gs.mean)
    print(f'2-D N={Nn} ok')

from mfs_b200.classical_filters_smoothers import brute_force_filter
from mfs_b200.functors import benes_drift, Dispersion, bernoulli_logistic_cubic
grid = np.linspace(-4., 4., 130)
for B3 in (3, 70):
    ysb = (np.random.default_rng(1).random((B3, 3)) < 0.5).astype(np.uint8)
    for method, steps in (('chapman-tme-3', 3), ('chapman-euler', 2), ('kolmogorov', 20)):
        brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ic.pdf(grid), grid, ysb, dt,
                           integration_steps=steps, pred_method=method, return_nell=True)
print('grid filter ok')
torch.cuda.synchronize()
print('ALL CASES RAN')

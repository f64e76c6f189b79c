"""One launch of the 2-D filter kernel on a seeded prey--predator batch (for ncu).  usage: nd_profile_case.py [N] [B] [T] [family]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mfs_b200.multi_dims.multi_indices import generate_graded_lexico_multi_indices as gen, gram_and_hankel_indices_graded_lexico as gh
from mfs_b200.multi_dims.filtering import moment_filter_nd_cms
from mfs_b200.multi_dims.moments import sde_cond_moments_euler_maruyama, sde_cond_moments_tme_normal, sde_cond_moments_tme
from mfs_b200.multi_dims.ss_models import prey_predator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
family = sys.argv[4] if len(sys.argv) > 4 else 'tme_normal'
mis = gen(2, 2 * N - 1); inds = gh(N, 2)
dt, _, ts, gs, drift, dispersion, emission, pmf, simulate = prey_predator(mis)
rng = np.random.Generator(np.random.PCG64(677))
_, xs, ys = simulate(rng, integration_steps=10, T=T, n=64)
ys = torch.from_numpy(np.tile(ys, (B // 64 + 1, 1))[:B].copy()).cuda()
if family == 'euler':
    fam, flag = sde_cond_moments_euler_maruyama(drift, dispersion, dt, mis), 'index'
elif family == 'tme':
    fam, flag = sde_cond_moments_tme(drift, dispersion, dt, 2), 'multi-index'
else:
    fam, flag = sde_cond_moments_tme_normal(drift, dispersion, dt, 2, mis), 'index'
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = moment_filter_nd_cms((fam[1], flag), fam[3], pmf, ys, (mis, inds), gs.cms, gs.mean, history='last', return_status=True)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'nd N={N} B={B} T={T} {family}: {ms:.2f} ms {B * T / ms * 1e3:.3e} steps/s diverged {(out[-1] >= 0).double().mean().item():.3f}')

"""BASELINE configs[2]: convergence sweep N = 2..15 (Hankel conditioning stress), 1e5 trajectories per N, fp64 tolerance
report against the C oracle (restatement of the reference's dense algorithm) and the Kalman filter.

OU + Gaussian likelihood with the exact Normal transition (dardel/convergence/convergence_mf.py:32-107, dt = 0.1, T = 100,
ell = 1, sigma = 0.5, R = 1), raw and central moments.  Writes a markdown table to stdout.
usage: python tools/convergence_report.py [B] [B_oracle]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mfs_b200 import synthetic
from mfs_b200.functors import gaussian
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
from mfs_b200.one_dim.moments import sde_cond_moments_normal_affine, raw_moment_of_normal, raw_to_central
from oracle import c_oracle as C
from oracle import mfs_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
BO = int(sys.argv[2]) if len(sys.argv) > 2 else 512
T, dt, ell, sigma = 100, 0.1, 1., 0.5
F, Sig = np.exp(-dt / ell), sigma ** 2 * (1 - np.exp(-2 * dt / ell))
ys = synthetic.ou_gaussian_ys_numpy(B, T, 669, dt, ell, sigma, 1.)
ys_d = torch.from_numpy(ys).cuda()
fam = sde_cond_moments_normal_affine(F, Sig)
kf = [O.kalman_filter_1d(F, Sig, 1., 1., 0., sigma ** 2, ys[k]) for k in range(64)]
kf_m, kf_v = np.stack([k[0] for k in kf]), np.stack([k[1] for k in kf])
print(f'| N | mode | steps/s | diverged GPU | diverged oracle ({BO}) | mean vs oracle (max abs) | var vs oracle (max rel) | '
      f'nell vs oracle (max rel) | mean vs Kalman (max abs) | var vs Kalman (max rel) |')
print('|---|---|---|---|---|---|---|---|---|---|')
for N in range(2, 16):
    rms0 = np.array([raw_moment_of_normal(0., sigma ** 2, p) for p in range(2 * N)])
    cms0 = raw_to_central(rms0)
    for mode in ('raw', 'central'):
        def run():
            if mode == 'raw':
                ms, nell, st = moment_filter_rms(fam[0], gaussian(1., 1.), rms0, ys_d, return_status=True)
                return ms, ms[..., 1], ms[..., 2] - ms[..., 1] ** 2, nell, st
            ms, mean, nell, st = moment_filter_cms(fam[1], fam[3], gaussian(1., 1.), cms0, 0., ys_d, return_status=True)
            return ms, mean, ms[..., 2], nell, st
        run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ms, mean, var, nell, st = run()
        e1.record()
        torch.cuda.synchronize()
        rate = B * T / (e0.elapsed_time(e1) * 1e-3)
        mean, var, nell, st = (t[:BO].cpu().numpy() for t in (mean, var, nell, st))
        div_gpu = float((ms[:, -1, 0] != ms[:, -1, 0]).double().mean().item())
        ref = C.filter_1d(mode, fam[0 if mode == 'raw' else 1], gaussian(1., 1.), rms0 if mode == 'raw' else cms0, ys[:BO],
                          mean0=None if mode == 'raw' else 0.)
        ref_mean = ref['ms'][..., 1] if mode == 'raw' else ref['mean']
        ref_var = ref['ms'][..., 2] - ref['ms'][..., 1] ** 2 if mode == 'raw' else ref['ms'][..., 2]
        ok = (st < 0) & (ref['status'] < 0)
        okk = ok[:64]
        fmt = lambda v: f'{v:.1e}'
        if ok.any():
            cols = [fmt(np.max(np.abs(mean[ok] - ref_mean[ok]))), fmt(np.max(np.abs(var[ok] / ref_var[ok] - 1))),
                    fmt(np.max(np.abs(nell[ok] / ref['nell'][ok] - 1)))]
        else:
            cols = ['-', '-', '-']
        if okk.any():
            cols += [fmt(np.max(np.abs(mean[:64][okk] - kf_m[okk]))), fmt(np.max(np.abs(var[:64][okk] / kf_v[okk] - 1)))]
        else:
            cols += ['-', '-']
        print(f'| {N} | {mode} | {rate:.2e} | {div_gpu:.4f} | {np.mean(ref["status"] >= 0):.4f} | ' + ' | '.join(cols) + ' |',
              flush=True)

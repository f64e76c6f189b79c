"""Throughput of the 1-D filter for N = 2..15 (the reference's time-profile experiment, dardel/time_profile/mf.py +
run_time_profile.sh:25: one trajectory per N on the CPU), here as filter-steps/s of a full batch on one GPU, with the
FP64 roofline fraction by the work model W(N) of SURVEY 8(d) / DESIGN 3 (atom-reuse recursion: W - 20 N (N + 1))."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200 import _lib
from mfs_b200.simulate import simulate_1d
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
from mfs_b200.one_dim.moments import sde_cond_moments_tme
from mfs_b200.one_dim.ss_models import benes_bernoulli

T = 100
peak, _ = _lib.fp64_peak(0, 4096)
print(f'# Time profile over N (Benes-Bernoulli, TME-3, T = {T}, history = none, one B200; FP64 peak measured {peak / 1e12:.1f} TFLOP/s)\n')
print('| N | filters | raw steps/s | central steps/s | diverged (raw) | W_reuse(N) flop/step | FP64 roofline frac (raw, all steps) |')
print('|---|---|---|---|---|---|---|')
for N in range(2, 16):
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    B = 148 * 128 * (48 if N <= 4 else 24 if N <= 8 else 12)
    _, _, ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 100 + N, scheme='benes_exact')
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    rates = []
    for mode in ('raw', 'central'):
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode == 'raw':
                out = moment_filter_rms(fam[0], pmf, ic.rms, ys, history='none', return_status=True)
            else:
                out = moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys, history='none', return_status=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best     # N = 2 runs 3 ms: host jitter shows
        rates.append(B * T / (best * 1e-3))
        if mode == 'raw':
            div = float((out[-1] >= 0).double().mean())
    W = (2 / 3) * N ** 3 + 85 * N ** 2 + 130 * N + 30 - 20 * N * (N + 1)
    print(f'| {N} | {B} | {rates[0]:.3e} | {rates[1]:.3e} | {div:.3f} | {W:.0f} | {rates[0] * W / peak:.3f} |')

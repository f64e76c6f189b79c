"""Headline case (Benes-Bernoulli raw, N = 8, T = 1000, 1e6 filters, full history) for several segment lengths of the
segmented execution with live-filter compaction (include/mfs_b200.h: segment_steps).  usage: python tools/segment_sweep.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200.one_dim.filtering import moment_filter_rms
from mfs_b200.one_dim.moments import sde_cond_moments_tme
from mfs_b200.one_dim.ss_models import benes_bernoulli
from mfs_b200.simulate import simulate_1d

N, B, T = 8, 1000000, 1000
dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
fam = sde_cond_moments_tme(drift, disp, dt, 3)
ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 667, scheme='benes_exact')[2]
bufs = {}
out = moment_filter_rms(fam[0], pmf, ic.rms, ys, history='full', return_status=True, out=bufs)
bufs.update({k: v for k, v in zip(('ms', 'nell', 'status'), out)})
for seg in (24, 32, 48, 64, 96, 128, 250):
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        moment_filter_rms(fam[0], pmf, ic.rms, ys, history='full', return_status=True, out=bufs, segment_steps=seg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    print(f'segment_steps={seg:4d}: {best:8.3f} ms  {B * T / best * 1e3:.4e} filter-steps/s', flush=True)

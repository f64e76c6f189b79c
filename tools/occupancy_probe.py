"""Raw Benes filter, T = 100, history none, for the given orders (same-box A/B of register caps).  usage: MFS_B200_LIB=... python tools/occupancy_probe.py tag N [N ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200.one_dim.filtering import moment_filter_rms
from mfs_b200.one_dim.moments import sde_cond_moments_tme
from mfs_b200.one_dim.ss_models import benes_bernoulli
from mfs_b200.simulate import simulate_1d

tag = sys.argv[1]
for N in [int(a) for a in sys.argv[2:]]:
    B, T = (909312 if N <= 4 else 454656 if N <= 8 else 227328), 100
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 100 + N, scheme='benes_exact')[2]
    best = None
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        moment_filter_rms(fam[0], pmf, ic.rms, ys, history='none', return_status=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    print(f'[{tag}] N={N}: {best:.3f} ms {B * T / best * 1e3:.4e} steps/s', flush=True)

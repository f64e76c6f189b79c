"""ys streaming A/B (north_star: "coalesced, vectorised loads"): the same records as a (B, T) row-major buffer (lane b
reads byte ys[b][t]: 32 sectors per warp load, each re-used for 32 steps from L1/L2) and as the transpose view of a
time-major (T, B) buffer (a warp reads 32 consecutive bytes: one sector per step).  N = 8 raw Benes TME-3.
usage: python tools/ys_layout_ab.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200.one_dim.filtering import moment_filter_rms
from mfs_b200.one_dim.moments import sde_cond_moments_tme
from mfs_b200.one_dim.ss_models import benes_bernoulli
from mfs_b200.simulate import simulate_1d

N = 8
dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
fam = sde_cond_moments_tme(drift, disp, dt, 3)
for B, T, history in ((454656, 100, 'none'), (1000000, 1000, 'meanvar'), (1000000, 1000, 'none')):
    ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 667, scheme='benes_exact')[2]
    ys_tm = ys.t().contiguous()            # (T, B) time-major
    res = {}
    for name, y in (('row-major (B,T)', ys), ('time-major (T,B) view', ys_tm.t())):
        bufs = {}
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = moment_filter_rms(fam[0], pmf, ic.rms, y, history=history, return_status=True, out=bufs)
            e1.record()
            torch.cuda.synchronize()
            bufs = {'nell': out[1], 'status': out[2]}
            if history == 'meanvar':
                bufs['meanvar'] = out[0]
        res[name] = (e0.elapsed_time(e1), out)
        print(f'N={N} B={B} T={T} {history} ys {name}: {res[name][0]:.3f} ms {B * T / res[name][0] * 1e3:.4e} steps/s', flush=True)
    a, b = res['row-major (B,T)'][1], res['time-major (T,B) view'][1]
    assert torch.equal(a[1].nan_to_num(-7.), b[1].nan_to_num(-7.)) and torch.equal(a[2], b[2]), 'layouts disagree'
    del ys, ys_tm, res, a, b
    torch.cuda.empty_cache()

"""One launch of the filter kernel on a seeded Benes--Bernoulli batch, for `ncu --set full`.
usage: python tools/profile_case.py [N] [B] [T] [mode] [history]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
from mfs_b200.one_dim.moments import sde_cond_moments_tme
from mfs_b200.one_dim.ss_models import benes_bernoulli
from mfs_b200.synthetic import benes_bernoulli_ys_torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
mode = sys.argv[4] if len(sys.argv) > 4 else 'raw'
recompute = os.environ.get('MFS_RECOMPUTE', '0') == '1'
history = sys.argv[5] if len(sys.argv) > 5 else 'full'
seg = int(os.environ['MFS_SEG']) if 'MFS_SEG' in os.environ else None
bufs = {}
dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
fam = sde_cond_moments_tme(drift, disp, dt, 3)
ys = benes_bernoulli_ys_torch(B, T, 667, 'cuda')
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if mode == 'raw':
        out = moment_filter_rms(fam[0], pmf, ic.rms, ys, history=history, return_status=True, recompute_predict_quadrature=recompute, segment_steps=seg, out=bufs)
        bufs = {'ms': out[0], 'nell': out[1], 'status': out[2]} if history != 'none' else {'nell': out[1], 'status': out[2]}
    else:
        out = moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys, history=history, return_status=True, recompute_predict_quadrature=recompute)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f'N={N} B={B} T={T} {mode} {history}: {ms:.3f} ms {B * T / ms * 1e3:.3e} steps/s diverged {(out[-1] >= 0).double().mean().item():.3f}')

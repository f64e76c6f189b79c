"""Fixed list of 1-D filter cases for same-box A/B runs of kernel builds (CUDA events, third of three runs).
usage: MFS_B200_LIB=path/to/lib.so python tools/ab_cases.py [tag] [--quick]

Cases: the headline (N=8 raw Benes TME-3, T=1000 full history, 1e6 filters), the same at the reference's horizon
T=100 without history, N=5 / N=12 / N=15, the central mode, and the estimation objective (well--Poisson, central,
TME-normal 2, N=7: the Normal-family + Poisson instance)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
from mfs_b200.one_dim.moments import sde_cond_moments_tme, sde_cond_moments_tme_normal
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson
from mfs_b200.simulate import simulate_1d

tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('--') else os.environ.get('MFS_B200_LIB', 'default')
quick = '--quick' in sys.argv
sweep = '--sweep' in sys.argv      # every order N = 2..15 at T = 100 (raw Benes) + the Normal / Poisson instance at a few N


def timed(fn, reps=3):
    ms = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    return ms, out


def benes(N, B, T, mode, history, literal=False):
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 667, scheme='benes_exact')[2]
    bufs = {}
    if mode == 'raw':
        fn = lambda: moment_filter_rms(fam[0], pmf, ic.rms, ys, history=history, return_status=True, out=bufs,
                                       recompute_predict_quadrature=literal)
    else:
        fn = lambda: moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys, history=history, return_status=True,
                                       out=bufs, recompute_predict_quadrature=literal)
    out = fn()
    keys = ('ms', 'nell', 'status') if mode == 'raw' else ('ms', 'mean', 'nell', 'status')
    bufs.update({k: v for k, v in zip(keys, out) if v is not None})
    if history == 'meanvar':
        bufs['meanvar'] = bufs.pop('ms')
    ms, out = timed(fn)
    st = out[-1]
    live = float(torch.where(st >= 0, st, torch.full_like(st, T)).double().sum().item()) / (B * T)
    print(f'[{tag}] benes N={N} B={B} T={T} {mode} {history}{" literal" if literal else ""}: {ms:.3f} ms '
          f'{B * T / ms * 1e3:.4e} steps/s (live {B * T * live / ms * 1e3:.4e}) diverged {(st >= 0).double().mean().item():.3f}',
          flush=True)
    del ys, bufs, out
    torch.cuda.empty_cache()


def well(N, B, T):
    dt, _, _, ic, drift, disp, _, pmf, _ = well_poisson(3., N)
    ys = simulate_1d(drift(3.), disp, dt, T, ic, pmf(3.), B, 7)[2]
    fam = sde_cond_moments_tme_normal(drift(3.), disp, dt, 2, N)
    fn = lambda: moment_filter_cms(fam[1], fam[3], pmf(3.), ic.cms, ic.mean, ys, history='none', return_status=True)
    ms, out = timed(fn)
    print(f'[{tag}] well-poisson N={N} B={B} T={T} central tme_normal2 none: {ms:.3f} ms {B * T / ms * 1e3:.4e} steps/s '
          f'diverged {(out[-1] >= 0).double().mean().item():.3f}', flush=True)


if sweep:
    for n in range(2, 16):
        benes(n, 909312 if n <= 4 else 454656 if n <= 8 else 227328, 100, 'raw', 'none')
    for n in (3, 5, 7, 9, 12):
        well(n, 75776 * 4 if n <= 7 else 75776 * 2, 200)
    sys.exit(0)
benes(8, 454656, 100, 'raw', 'none')
benes(8, 1000000, 1000, 'raw', 'full')
well(7, 75776 * 4, 1000)
if not quick:
    benes(8, 454656, 100, 'raw', 'full')
    benes(8, 454656, 100, 'raw', 'none', literal=True)
    benes(8, 454656, 100, 'central', 'none')
    benes(5, 454656 * 2, 100, 'raw', 'none')
    benes(12, 227328, 100, 'central', 'none')
    benes(15, 113664, 100, 'raw', 'none')
    benes(3, 909312, 100, 'raw', 'none')
    benes(8, 1000000, 1000, 'raw', 'meanvar')

"""Distance to exact arithmetic, per N = 2..8 (VERDICT r1 "weak #1"): tests/golden/golden_exact_1d.npz holds the reference's
Benes--Bernoulli recursion in 60-digit arithmetic on 12 records per N.  Prints, as a markdown table, the per-record max
relative moment error of the NumPy/LAPACK oracle, the C oracle and the CUDA path (default atom-reuse recursion and the
literal two-quadrature recursion) against it, and the paired ratio CUDA / max(CPU).
usage: python tools/exactness_report.py > profiles/r2_exactness_report.md   (needs a GPU)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from test_oracle_mp import cpu_scores, moment_err
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
from mfs_b200.one_dim.moments import sde_cond_moments_tme
from mfs_b200.one_dim.ss_models import benes_bernoulli

print('# Round 2 -- distance of every fp64 implementation to exact arithmetic (Benes-Bernoulli, TME-3, T = 100, 12 records per N)\n')
print('Exact = the reference recursion evaluated with 60 significant digits (`oracle/mfs_oracle_mp.py`, fixture '
      '`tests/golden/golden_exact_1d.npz`).  Entries: per-record max over (t, p) of the relative moment error (per-order floor '
      '(E X^2)^(p/2)); median / max over the records.  "paired" = per-record error(CUDA) / max(error(LAPACK), error(C)).\n')
print('| N | mode | LAPACK oracle | C oracle | CUDA (default) | CUDA (literal two quadratures) | paired ratio median / geo-mean | nell: CUDA max rel err |')
print('|---|---|---|---|---|---|---|---|')
fmt = lambda e: f'{np.median(e):.1e} / {e.max():.1e}'
for N in range(2, 9):
    dt, _, ts, ic, drift, disp, logistic, pmf, _ = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    g, ok, e_c, e_l, _ = cpu_scores(N, 'raw')
    ys = torch.from_numpy(g[f'N{N}/ys']).cuda()
    res = {}
    for literal in (False, True):
        rmss, nell = moment_filter_rms(fam[0], pmf, g[f'N{N}/rms0'], ys, recompute_predict_quadrature=literal)
        res[literal] = moment_err(rmss.cpu().numpy()[ok], g[f'N{N}/rmss'][ok])
        if not literal:
            n_err = np.abs(nell.cpu().numpy()[ok] - g[f'N{N}/nell_raw'][ok]) / np.abs(g[f'N{N}/nell_raw'][ok])
    ratio = res[False] / np.maximum(e_c, e_l)
    print(f'| {N} | raw | {fmt(e_l)} | {fmt(e_c)} | {fmt(res[False])} | {fmt(res[True])} | {np.median(ratio):.2f} / '
          f'{np.exp(np.mean(np.log(ratio))):.2f} | {n_err.max():.1e} |')
    g, okc, e_cc, _, _ = cpu_scores(N, 'central')
    cmss, means, nell_c = moment_filter_cms(fam[1], fam[3], pmf, g[f'N{N}/cms0'], float(g[f'N{N}/mean0']), ys)
    e_gc = moment_err(cmss.cpu().numpy()[okc], g[f'N{N}/cmss'][okc])
    n_err = np.abs(nell_c.cpu().numpy()[okc] - g[f'N{N}/nell_central'][okc]) / np.abs(g[f'N{N}/nell_central'][okc])
    ratio = e_gc / e_cc
    print(f'| {N} | central | -- | {fmt(e_cc)} | {fmt(e_gc)} | -- | {np.median(ratio):.2f} / {np.exp(np.mean(np.log(ratio))):.2f} | {n_err.max():.1e} |')

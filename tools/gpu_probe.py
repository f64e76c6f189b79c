"""Quick on-box probe: FP64 peak, parity against golden fixtures, first throughput numbers."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mfs_b200
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms, moment_filter_scms
from mfs_b200.one_dim.moments import sde_cond_moments_tme, sde_cond_moments_euler, sde_cond_moments_tme_normal
from mfs_b200.one_dim.ss_models import benes_bernoulli

print('fp64 peak', mfs_b200.fp64_peak(0, 4096))
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
for N in (5, 8):
    g = np.load(os.path.join(G, f'golden_filter_1d_benes_N{N}.npz'))
    dt, T, ts, ic, drift, disp, logistic, pmf, sim = benes_bernoulli(N)
    ys = torch.from_numpy(g['ys']).cuda()
    fams = {'tme3': sde_cond_moments_tme(drift, disp, dt, 3), 'tme2': sde_cond_moments_tme(drift, disp, dt, 2),
            'euler': sde_cond_moments_euler(drift, disp, dt, N), 'tme_normal3': sde_cond_moments_tme_normal(drift, disp, dt, 3, N)}
    for name, fam in fams.items():
        rmss, nell, st = moment_filter_rms(fam[0], pmf, ic.rms, ys, return_status=True)
        rmss = rmss.cpu().numpy(); nell = nell.cpu().numpy()
        errs = []
        for k in range(ys.shape[0]):
            ref = g[f'{name}/rms/{k}/rmss']
            errs.append(np.nanmax(np.abs(rmss[k] / ref - 1)))
        print(N, name, 'raw maxrel', ['%.1e' % e for e in errs], 'nell err', ['%.1e' % abs(nell[k] - float(g[f'{name}/rms/{k}/nell'])) for k in range(4)], st.cpu().numpy())
        cmss, means, nell = moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys)
        cmss = cmss.cpu().numpy(); means = means.cpu().numpy(); nell = nell.cpu().numpy()
        errs = [np.nanmax(np.abs(cmss[k][:, 2:] / g[f'{name}/cms/{k}/cmss'][:, 2:] - 1)) for k in range(4)]
        print(N, name, 'cms maxrel', ['%.1e' % e for e in errs], 'mean err', ['%.1e' % np.max(np.abs(means[k] - g[f'{name}/cms/{k}/means'])) for k in range(4)],
              'nell err', ['%.1e' % abs(nell[k] - float(g[f'{name}/cms/{k}/nell'])) for k in range(4)])
        if name in ('tme2', 'tme3'):
            s, means, scales, nell = moment_filter_scms(fam[2], fam[4], pmf, ic.scms, ic.mean, np.sqrt(ic.variance), ys)
            s = s.cpu().numpy()
            errs = [np.nanmax(np.abs(s[k][:, 3:] / g[f'{name}/scms/{k}/scmss'][:, 3:] - 1)) for k in range(4)]
            print(N, name, 'scms maxrel', ['%.1e' % e for e in errs], 'nell err', ['%.1e' % abs(nell[k].item() - float(g[f'{name}/scms/{k}/nell'])) for k in range(4)])

# throughput
rng = np.random.default_rng(0)
for N in (5, 8):
    dt, T, ts, ic, drift, disp, logistic, pmf, sim = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    for (B, T, hist) in ((148 * 128 * 12, 100, 'full'), (148 * 128 * 12, 100, 'none'), (148 * 128 * 12, 1000, 'none')):
        x0 = ic.sampler(rng, B)
        xs = sim(x0, rng, T)
        ys = torch.from_numpy((rng.random((T, B)) < logistic(xs)).T.astype(np.uint8).copy()).cuda()
        for mode in ('raw', 'central'):
            def run():
                if mode == 'raw':
                    return moment_filter_rms(fam[0], pmf, ic.rms, ys, history=hist, return_status=True)
                return moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys, history=hist, return_status=True)
            out = run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = out[-1].cpu().numpy()
            live = np.where(st < 0, T, st).sum() / (B * T)
            print(f'N={N} B={B} T={T} {mode} {hist}: {ms:.2f} ms  {B*T/ms*1e3:.3e} steps/s  diverged {np.mean(st>=0):.3f} live-frac {live:.3f}')

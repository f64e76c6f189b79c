"""Throughput of the device simulators (CUDA events, one warm-up): trajectory-steps/s and sub-steps/s."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfs_b200.simulate import simulate_1d, simulate_prey_predator
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson
from mfs_b200.multi_dims.ss_models import prey_predator
from mfs_b200.multi_dims import multi_indices as MI


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dt, T, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(8)
for B, T_, steps, scheme in ((1000000, 100, 100, 'tme'), (1000000, 1000, 1, 'benes_exact'), (1000000, 1000, 10, 'tme')):
    ms = timed(lambda: simulate_1d(drift, disp, dt, T_, ic, pmf, B, 1, integration_steps=steps, scheme=scheme))
    print(f'benes-bernoulli {scheme:12s} B={B} T={T_} sub-steps={steps}: {ms:9.3f} ms  {B * T_ / ms * 1e3:.3e} traj-steps/s  '
          f'{B * T_ * steps / ms * 1e3:.3e} sub-steps/s')
dtw, Tw, _, icw, driftw, dispw, _, pmfw, _ = well_poisson(3., 7)
ms = timed(lambda: simulate_1d(driftw(3.), dispw, dtw, Tw, icw, pmfw(3.), 100000, 2))
print(f'well-poisson tme-3 B=100000 T={Tw} sub-steps=100: {ms:9.3f} ms  {100000 * Tw / ms * 1e3:.3e} traj-steps/s  {100000 * Tw * 100 / ms * 1e3:.3e} sub-steps/s')
mi = MI.generate_graded_lexico_multi_indices(2, 3, 0)
dtl, Tl, _, gs, driftl, displ, _, pmfl, _ = prey_predator(mi)
ms = timed(lambda: simulate_prey_predator(driftl, displ, dtl, Tl, gs, pmfl, 100000, 3))
print(f'prey-predator milstein B=100000 T={Tl} sub-steps=100: {ms:9.3f} ms  {100000 * Tl / ms * 1e3:.3e} traj-steps/s  {100000 * Tl * 100 / ms * 1e3:.3e} sub-steps/s')

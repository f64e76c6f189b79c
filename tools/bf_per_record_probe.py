"""Grid filter with one grid per record (the reference's call shape, dardel/benes_bernoulli/brute_force.py:73-76) at the
paper's setting n = 2000, 100 sub-steps: time per record and time step, against the L2 -> SM bound of re-reading the 32 MB
operator once per sub-step.  usage: python tools/bf_per_record_probe.py [records] [T]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mfs_b200.classical_filters_smoothers import brute_force_filter
from mfs_b200.functors import benes_drift, Dispersion, bernoulli_logistic_cubic
from mfs_b200.one_dim.ss_models import benes_bernoulli
from mfs_b200.simulate import simulate_1d

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, steps = 2000, 100
dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(8)
ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 678)[2]
rng = np.random.default_rng(3)
xs = np.stack([np.linspace(-6. - rng.random(), 6. + rng.random(), n) for _ in range(B)])
ip = np.stack([ic.pdf(xs[k]) for k in range(B)])
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, nell = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys, dt,
                                   integration_steps=steps, pred_method='chapman-tme-3', history='last', return_nell=True)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
sub = B * T * steps
print(f'per-record grids: {B} records x T={T} x {steps} sub-steps, n={n}: {ms:.1f} ms = {ms / (B * T):.2f} ms per record time-step, '
      f'{ms * 1e3 / sub:.2f} us per sub-step (one 2000 x 2000 matrix-vector product on an L2-resident 32 MB operator: '
      f'{32e6 / (ms * 1e-3 / sub) / 1e12:.2f} TB/s of operator reads, {2 * n * n / (ms * 1e-3 / sub) / 1e12:.3f} TFLOP/s)')
# the same records with the operator's 100th power formed once per record (power_operator=True)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out_p, nell_p = brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs, ys, dt,
                                       integration_steps=steps, pred_method='chapman-tme-3', history='last', return_nell=True,
                                       power_operator=True)
    e1.record()
    torch.cuda.synchronize()
ms_p = e0.elapsed_time(e1)
rel = float(((out_p - out).abs() / out.abs().amax(dim=-1, keepdim=True)).max())
print(f'power_operator=True: {ms_p:.1f} ms = {ms_p / B:.2f} ms per record ({ms / ms_p:.1f}x faster than the literal recursion at '
      f'T={T}); max relative density difference {rel:.1e}, nell {float(((nell_p - nell).abs() / nell.abs()).max()):.1e}')

"""GPU probe of the brute-force grid filter: DMMA / DFMA peaks and GEMM sub-step throughput (CUDA events)."""
import sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mfs_b200
from mfs_b200 import _lib
from mfs_b200.classical_filters_smoothers import brute_force_filter
from mfs_b200.functors import benes_drift, Dispersion, bernoulli_logistic_cubic

fp64, _ = _lib.fp64_peak(0, 4096)
dmma, _ = _lib.dmma_peak(0, 4096)
print(f'DFMA peak {fp64/1e12:.2f} TF  DMMA m8n8k4 peak {dmma/1e12:.2f} TF')
n = int(os.environ.get('BF_N', 2000))
for B in [int(b) for b in os.environ.get('BF_B', '1024,16384,65536').split(',')]:
    T, steps = 2, 50
    rng = np.random.Generator(np.random.PCG64(1))
    xs = np.linspace(-6., 6., n)
    ys = torch.from_numpy((rng.random((B, T)) < 0.5).astype(np.uint8)).cuda()
    ip = 0.5 * np.exp(-0.5 * (xs + 0.5) ** 2 / 0.05) / math.sqrt(2 * math.pi * 0.05) + 0.5 * np.exp(-0.5 * (xs - 0.5) ** 2 / 0.05) / math.sqrt(2 * math.pi * 0.05)
    args = (benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs)
    out = brute_force_filter(*args, ys, 1e-2, integration_steps=steps, pred_method='chapman-tme-3', history='last')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = brute_force_filter(*args, ys, 1e-2, integration_steps=steps, pred_method='chapman-tme-3', history='last')
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flop = 2.0 * n * n * B * steps * T
    print(f'n={n} B={B}: {ms:.1f} ms for {T}x{steps} sub-steps -> {flop/ms/1e9:.2f} TFLOP/s ({flop/ms/1e9/(dmma/1e12)*100:.1f}% of DMMA peak); mass {float(torch.trapezoid(out[0], torch.from_numpy(xs).cuda())):.15f}')

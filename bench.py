#!/usr/bin/env python
"""Headline benchmark: Benes--Bernoulli moment filter, N=8 nodes (16 raw moments), T=1000, 1e6 independent filters per
GPU (BASELINE.json configs[1]); metric = filter-steps/s (batch x T / time), fp64.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 3 --warmup 3
    python bench.py --impl reference --steps 2 --warmup 1       # CPU arm: the C restatement of the reference algorithm

One "step" = one pass of the filter over the whole batch (B x T filter-steps; T = 1000 runs as 16 segment launches with
live-filter compaction + one NaN-tail launch per GPU).
  value   inputs (ys uint8, 1 GB) resident in HBM, full moment history (B, T, 16) written to HBM, CUDA events on the
          launching stream, barrier + synchronize on both sides, max over ranks.  ys (1 GB) > L2 (126 MB).
  e2e     the public API `moment_filter_rms(...)` with HOST buffers (pinned), H2D of ys, kernel, D2H of the full
          history inside the timed region, weak-scaled (fixed filters per GPU); `e2e_modes` reports the same call for the
          three output modes full / meanvar (mean, variance per step) / none (nell only).
  roofline  FP64 FMA pipe (the path is compute-bound: ~53 flop/B): algorithmic flops W(8) per filter-step
          x LIVE filter-steps / kernel time, against the FP64 FMA peak measured live by mfs_fp64_peak (MEASURED_PEAKS.json
          holds no FP64 figure).  `hbm` sub-object gives the secondary HBM figure against MEASURED_PEAKS.json.
  secondary  the other BASELINE configurations at their stated sizes, sharded over the ranks (1/8 of the stated size per
          rank): theta grid + argmin over NCCL (configs[3]), 2-D prey--predator filter and brute-force grid filter
          (configs[4]); on one GPU also the data simulator and the nell + gradient kernel.
  cpu_baseline  oracle/libmfs_oracle.so (C restatement of the reference's dense algorithm, pthreads over filters) on a
          bounded sample, all host cores; `cpu_baseline_torch`: the same algorithm as batched torch-CPU linalg calls
          (baseline/torch_cpu_filter.py).  The JAX reference itself cannot run here (no jax / tme, no network).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'benes_bernoulli_filter_steps_per_s'
UNIT = 'filter-steps/s'


def work_per_step(N: int, two_quadratures: bool) -> float:
    """Algorithmic FP64 flop per filter-step.  SURVEY.md 8(d): W(N) = 2 W_quad + W_pred + W_upd
    = 2/3 N^3 + 85 N^2 + 130 N + 30 for the literal recursion (two moment quadratures per step), with
    W_quad = N^3/3 + N^2/2 + 4N [moments -> Jacobi coefficients] + 20 N (N+1) [tridiagonal QL].  The default kernel
    re-uses the posterior atoms as the prediction quadrature, i.e. it runs ONE QL per step (the second quadrature
    keeps only its Jacobi/pivot part): W_reuse(N) = W(N) - 20 N (N+1).  Only flops actually required are claimed."""
    w = 2. / 3. * N ** 3 + 85. * N ** 2 + 130. * N + 30.
    return w if two_quadratures else w - 20. * N * (N + 1)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7),
                              ('sw_power_cap', 8)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(mx)] if sm else []
        return {'sm_mhz': float(np.median(busy or sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs next to its GPU (NVML's ideal affinity) while the pinned host buffers of the e2e
    legs are allocated and filled, so that on a two-socket host they land on the GPU's own NUMA node instead of wherever
    the rank happened to be scheduled.  Returns (previous affinity, number of CPUs bound to) or (None, 0) if NVML or the
    affinity call is unavailable; the caller restores the previous affinity afterwards."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(props.uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        prev = os.sched_getaffinity(0)
        local = cpus & prev
        if not local or local == prev:
            return None, len(prev)
        os.sched_setaffinity(0, local)
        return prev, len(local)
    except Exception:
        return None, 0


def cpu_baseline(N, T, target_seconds=15., threads=0, seed=666 + 1):
    """Time the C oracle on a bounded sample of the same workload.  Returns (steps/s, cores, sample string)."""
    from oracle import c_oracle
    from mfs_b200.one_dim.moments import sde_cond_moments_tme
    from mfs_b200.one_dim.ss_models import benes_bernoulli
    from mfs_b200.synthetic import benes_bernoulli_ys_numpy
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    cores = c_oracle.max_threads() if threads <= 0 else threads
    probe_b = 16 * cores
    ys = benes_bernoulli_ys_numpy(probe_b, T, seed)
    t0 = time.perf_counter()
    c_oracle.filter_1d('raw', fam[0], pmf, ic.rms, ys, history='full', num_threads=cores)
    rate = probe_b * T / (time.perf_counter() - t0)
    b = int(max(probe_b, min(262144, rate * target_seconds / T)) // (16 * cores) * (16 * cores))
    ys = benes_bernoulli_ys_numpy(b, T, seed + 1)
    t0 = time.perf_counter()
    out = c_oracle.filter_1d('raw', fam[0], pmf, ic.rms, ys, history='full', num_threads=cores)
    el = time.perf_counter() - t0
    return b * T / el, int(out['threads']), f'{b} filters x T={T} (first {b} of the seeded workload), full history, {el:.1f} s', b


def secondary_legs(rank, world, device_index, fp64_peak, scale=1.0, single_gpu_extras=True):
    """The other BASELINE configurations of the path, at their STATED sizes, sharded over the ranks (weak scaling: every
    rank owns 1/8 of the stated size, so the 8-GPU job is exactly the stated configuration and `--gpus 1` runs one such
    shard).  CUDA events, barrier on both sides, max over ranks.

      theta_grid   configs[3]: nell over a 512 x 512 theta grid x (125 x ranks) records, well--Poisson, central moments,
                   TME-normal order 2, N = 7, T = 1000 (dardel/parameter_estimation/mf.py:37-73); the flattened theta
                   grid is sharded over the ranks and the per-record argmin is combined over NCCL.
      nd_filter    configs[4]: prey--predator 2-D filter, N = 5, central, tme_normal_2 and tme_2, T = 2000, 12 500
                   records per rank (dardel/run_prey_predator_mf.sh:29-30).
      grid_filter  configs[4]: brute-force grid filter on Benes--Bernoulli, n = 2000, 100 sub-steps, T = 100, 12 500
                   records per rank on one grid (dardel/benes_bernoulli/brute_force.py:21-28): FP64 DMMA GEMMs.
    `scale` < 1 shrinks the record counts (smoke runs); the JSON states the sizes that ran."""
    import math
    import torch
    import torch.distributed as dist
    from mfs_b200 import _lib
    from mfs_b200.parallel import shard_bounds, local_argmin, running_argmin, argmin_over_shards, gather_filters
    out = {}
    dev = torch.device('cuda', device_index)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[device_index])
        torch.cuda.synchronize(dev)

    def timed(fn, warm=None):
        (warm or fn)()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        r = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), r

    def total(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return float(t[0])

    # ---- configs[3]: theta grid + argmin over NCCL -------------------------------------------------------------------
    from mfs_b200.one_dim.filtering import moment_filter_cms
    from mfs_b200.one_dim.moments import sde_cond_moments_tme_normal as tme_normal_1d
    from mfs_b200.one_dim.ss_models import well_poisson
    from mfs_b200.simulate import simulate_1d
    Ng, G, Tw = 7, 512, 1000
    n_traj = max(8, int(round(125 * scale))) * world
    dtw, _, _, icw, driftw, dispw, _, pmfw, _ = well_poisson(3., Ng)
    ysw = simulate_1d(driftw(3.), dispw, dtw, Tw, icw, pmfw(3.), n_traj, 670, device=dev)[2]   # same records on every rank
    th1, th2 = np.meshgrid(np.linspace(0.5, 6., G), np.linspace(0.5, 6., G), indexing='ij')
    th1, th2 = th1.reshape(-1), th2.reshape(-1)
    lo, hi = shard_bounds(G * G, rank, world)
    chunk = max(1024, (1 << 23) // n_traj)          # ~8M filters per call

    def theta_pass(lo_, hi_, keep_cols=0):
        best, kept, bad = None, [], 0.
        for c0 in range(lo_, hi_, chunk):
            c1 = min(hi_, c0 + chunk)
            fam = tme_normal_1d(driftw(th1[c0:c1, None]), dispw, dtw, 2, Ng)
            _, _, nell, st = moment_filter_cms(fam[1], fam[3], pmfw(th2[c0:c1, None]), icw.cms, icw.mean, ysw,
                                               history='none', return_status=True)
            best = running_argmin(best, nell, theta_offset=c0)
            bad += float((st >= 0).sum())
            if keep_cols:
                kept.append(nell[:, :keep_cols].clone())
        v, a = argmin_over_shards(*best)
        return v, a, bad, (torch.cat(kept) if keep_cols else None)

    ms, (val, arg, bad, _) = timed(lambda: theta_pass(lo, hi), warm=lambda: theta_pass(lo, min(hi, lo + chunk)))
    # cross-check of the collective (untimed): argmin of the all-gathered nell table for the first records
    cols = min(4, n_traj)
    _, _, _, part = theta_pass(lo, hi, keep_cols=cols)
    full = gather_filters(part, G * G)
    v_f, a_f = local_argmin(full)
    argmin_ok = bool(torch.equal(a_f, arg[:cols]) and torch.equal(v_f, val[:cols]))
    est = np.stack([th1[arg.cpu().numpy()], th2[arg.cpu().numpy()]], axis=1)
    n_filters = G * G * n_traj
    out['theta_grid'] = {
        'metric': 'well_poisson_theta_grid_filter_steps_per_s', 'value': n_filters * Tw / (ms * 1e-3), 'unit': UNIT,
        'config': f'BASELINE configs[3]: {G} x {G} theta grid over [0.5, 6]^2 x {n_traj} records (125 per rank), '
                  f'well-Poisson, central moments, TME-normal order 2, N={Ng}, T={Tw}; theta grid sharded over {world} rank(s), '
                  f'per-record (min, argmin) combined by all-gather over {"NCCL" if world > 1 else "one process"}',
        'filters': n_filters, 'ms': ms, 'diverged_frac': total(bad) / n_filters,
        'argmin_matches_gathered_table': argmin_ok,
        'median_theta_hat': [float(np.median(est[:, 0])), float(np.median(est[:, 1]))],
        'mean_abs_error_theta_hat': [float(np.mean(np.abs(est[:, 0] - 3.))), float(np.mean(np.abs(est[:, 1] - 3.)))]}
    del ysw, full, part
    torch.cuda.empty_cache()

    # ---- configs[4] (a): 2-D prey--predator moment filter -------------------------------------------------------------
    from mfs_b200.multi_dims.multi_indices import (generate_graded_lexico_multi_indices,
                                                   gram_and_hankel_indices_graded_lexico)
    from mfs_b200.multi_dims.filtering import moment_filter_nd_cms
    from mfs_b200.multi_dims.moments import sde_cond_moments_tme_normal, sde_cond_moments_tme
    from mfs_b200.multi_dims.ss_models import prey_predator
    from mfs_b200.simulate import simulate_prey_predator
    N, T2 = 5, 2000
    B2 = max(512, int(round(12500 * scale)))
    mis = generate_graded_lexico_multi_indices(2, 2 * N - 1, 0)
    inds = gram_and_hankel_indices_graded_lexico(N, 2)
    dt2, _, _, gs, drift2, disp2, _, pmf2, _ = prey_predator(mis)
    ys2 = simulate_prey_predator(drift2, disp2, dt2, T2, gs, pmf2, B2, 677, traj_offset=rank * B2, device=dev)[2]
    s, z = N * (N + 1) // 2, N * (2 * N + 1)
    # SURVEY.md 8(d) nominal work model of one 2-D filter step (d = 2): two quadratures + transition + update
    w_quad = s ** 3 / 3 + 2 * 2 * s ** 3 + 9 * 2 * s ** 3 + s ** 2 * (2 * s + 2)
    w_nd = 2 * w_quad + s ** 2 * (80 + 6 * z) + s ** 2 * (45 + 4 * z)
    nd = {}
    for name, flag, fam in (('tme_normal_2', 'index', sde_cond_moments_tme_normal(drift2, disp2, dt2, 2, mis)),
                            ('tme_2', 'multi-index', sde_cond_moments_tme(drift2, disp2, dt2, 2))):
        ms, res = timed(lambda: moment_filter_nd_cms((fam[1], flag), fam[3], pmf2, ys2, (mis, inds), gs.cms, gs.mean,
                                                      history='last', return_status=True),
                        warm=lambda: moment_filter_nd_cms((fam[1], flag), fam[3], pmf2, ys2[:512, :64].contiguous(),
                                                          (mis, inds), gs.cms, gs.mean, history='last'))
        rate = world * B2 * T2 / (ms * 1e-3)
        nd[name] = {'value': rate, 'kernel_ms': ms, 'fp64_frac_nominal': rate / world * w_nd / fp64_peak,
                    'diverged_frac': total(float((res[-1] >= 0).sum())) / (world * B2)}
    out['nd_filter'] = {'metric': 'prey_predator_2d_filter_steps_per_s', 'value': nd['tme_normal_2']['value'], 'unit': UNIT,
                        'config': f'BASELINE configs[4]: prey-predator 2-D filter, N={N} (z={z} moments, {s * s} nodes), central '
                                  f'moments, T={T2}, {B2} records per rank x {world} rank(s) (Milstein-simulated on the device, '
                                  f'100 sub-steps), history=last', 'nominal_flop_per_step': w_nd, 'transitions': nd}
    del ys2
    torch.cuda.empty_cache()

    # ---- configs[4] (b): brute-force grid filter --------------------------------------------------------------------
    from mfs_b200.classical_filters_smoothers import brute_force_filter
    from mfs_b200.functors import benes_drift, Dispersion, bernoulli_logistic_cubic
    from mfs_b200.one_dim.ss_models import benes_bernoulli
    n, steps, Tg = 2000, 100, 100
    Bg = max(256, int(round(12500 * scale)))
    dtb, _, _, icb, driftb, dispb, _, pmfb, _ = benes_bernoulli(8)
    ysg = simulate_1d(driftb, dispb, dtb, Tg, icb, pmfb, Bg, 678, traj_offset=rank * Bg, device=dev)[2]
    xs = np.linspace(-6., 6., n)
    ip = icb.pdf(xs)
    ms, (pdf, nell_g) = timed(lambda: brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip, xs,
                                                         ysg, dtb, integration_steps=steps, pred_method='chapman-tme-3',
                                                         history='last', return_nell=True),
                              warm=lambda: brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.),
                                                              ip, xs, ysg[:, :1].contiguous(), dtb, integration_steps=steps,
                                                              pred_method='chapman-tme-3', history='last', return_nell=True))
    dmma, _ = _lib.dmma_peak(device_index, 4096)
    flop = 2. * n * n * Bg * steps * Tg
    mass = torch.trapezoid(pdf, torch.from_numpy(xs).to(dev), dim=-1)
    out['grid_filter'] = {'metric': 'brute_force_filter_time_steps_per_s', 'value': world * Bg * Tg / (ms * 1e-3),
                          'unit': 'filter time-steps/s (each = 100 sub-step GEMMs on a 2000-point grid)',
                          'config': f'BASELINE configs[4]: n_grid={n}, {steps} integration sub-steps, chapman-tme-3, T={Tg}, '
                                    f'{Bg} records per rank x {world} rank(s) on one grid',
                          'ms': ms, 'tflops_per_gpu': flop / (ms * 1e-3) / 1e12,
                          'roofline': {'bound': 'fp64_tensor', 'achieved': flop / (ms * 1e-3) / 1e12,
                                       'peak': dmma / 1e12, 'unit': 'TFLOP/s', 'frac': flop / (ms * 1e-3) / dmma,
                                       'peak_source': 'measured live: mfs_dmma_peak (DMMA m8n8k4 micro-benchmark)'},
                          'max_abs_mass_error': float((mass - 1.).abs().max()),
                          'finite_nell_frac': total(float(torch.isfinite(nell_g).sum())) / (world * Bg)}
    # the same job with the operator's 100th power formed once (power_operator=True: one contraction per time step)
    ms_p, (pdf_p, nell_p) = timed(lambda: brute_force_filter(benes_drift(), Dispersion(1.), bernoulli_logistic_cubic(5., 0.), ip,
                                                             xs, ysg, dtb, integration_steps=steps, pred_method='chapman-tme-3',
                                                             history='last', return_nell=True, power_operator=True))
    scale_p = pdf.abs().amax(dim=-1, keepdim=True)
    out['grid_filter']['power_operator'] = {
        'value': world * Bg * Tg / (ms_p * 1e-3), 'ms': ms_p, 'speedup_over_literal': ms / ms_p,
        'max_rel_density_diff_vs_literal': float(((pdf_p - pdf).abs() / scale_p).max()),
        'max_rel_nell_diff_vs_literal': float(((nell_p - nell_g).abs() / nell_g.abs()).max()),
        'note': 'opt-in (MFS_BF_FLAG_POWER_OPERATOR): the sub-steps of a time step apply the same linear operator, so its '
                f'{steps}th power is formed once per call by binary powering on the DMMA GEMM and every time step is one '
                'contraction; same mathematics, different order of summation'}
    del ysg, pdf, pdf_p
    torch.cuda.empty_cache()
    if not (single_gpu_extras and world == 1):
        return out

    # ---- the reference's data simulator (ss_models.py:49-54: 100 TME-3 Gaussian sub-steps per dt, T = 100) + Bernoulli
    Bs, Tb = 1000000, 100
    ms, _ = timed(lambda: simulate_1d(driftb, dispb, dtb, Tb, icb, pmfb, Bs, 1, device=dev))
    out['simulator'] = {'metric': 'benes_bernoulli_simulated_trajectory_steps_per_s', 'value': Bs * Tb / (ms * 1e-3),
                        'unit': 'trajectory-steps/s (each = 100 TME-3 Gaussian sub-steps + one Bernoulli draw)',
                        'config': f'{Bs} trajectories x T={Tb}, 100 sub-steps, Philox4x32-10 + Box-Muller in the kernel',
                        'kernel_ms': ms, 'sub_steps_per_s': Bs * Tb * 100 / (ms * 1e-3)}
    # ---- nell + gradient in one pass (the objective of dardel/parameter_estimation/mf.py:37-54: well--Poisson, central
    #      moments, TME-normal order 2, N = 7, T = 1000), next to the value-only kernel on the same records
    from mfs_b200.one_dim.gradients import moment_filter_cms_value_and_grad
    Bq = 148 * 64 * 8
    ysw = simulate_1d(driftw(3.), dispw, dtw, Tw, icw, pmfw(3.), Bq, 7, device=dev)[2]
    famw = tme_normal_1d(driftw(3.), dispw, dtw, 2, Ng)
    ms_v, _ = timed(lambda: moment_filter_cms(famw[1], famw[3], pmfw(3.), icw.cms, icw.mean, ysw, history='none'))
    ms_g, rg = timed(lambda: moment_filter_cms_value_and_grad(famw[1], famw[3], pmfw(3.), icw.cms, icw.mean, ysw))
    out['gradient'] = {'metric': 'well_poisson_nell_and_gradient_filter_steps_per_s', 'value': Bq * Tw / (ms_g * 1e-3),
                       'unit': UNIT, 'config': f'N={Ng}, central, TME-normal order 2, {Bq} filters x T={Tw}, nell + d nell / d(theta1, theta2) '
                                               f'by forward-mode duals in one kernel', 'kernel_ms': ms_g,
                       'value_only_kernel_ms': ms_v, 'value_only_steps_per_s': Bq * Tw / (ms_v * 1e-3),
                       'cost_vs_value_only': ms_g / ms_v,
                       'finite_frac': float(torch.isfinite(rg[1]).all(dim=1).double().mean().item())}
    return out


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The JAX reference cannot be installed or run
    in this image (jax, jaxlib, tme absent; no network), so this arm times the C restatement of its algorithm
    (oracle/mfs_oracle.c) with all host threads -- rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    N, T = args.N, args.T
    from oracle import c_oracle
    from mfs_b200.one_dim.moments import sde_cond_moments_tme
    from mfs_b200.one_dim.ss_models import benes_bernoulli
    from mfs_b200.synthetic import benes_bernoulli_ys_numpy
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)
    cores = c_oracle.max_threads()
    # size one step for ~ref_step_seconds of CPU work
    probe_b = 16 * cores
    ys = benes_bernoulli_ys_numpy(probe_b, T, 667)
    t0 = time.perf_counter()
    c_oracle.filter_1d('raw', fam[0], pmf, ic.rms, ys, history='full', num_threads=cores)
    rate = probe_b * T / (time.perf_counter() - t0)
    b = int(max(probe_b, rate * args.ref_step_seconds / T) // (16 * cores) * (16 * cores))
    ys = benes_bernoulli_ys_numpy(b, T, 668)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = c_oracle.filter_1d('raw', fam[0], pmf, ic.rms, ys, history='full', num_threads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = args.steps * b * T / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'Benes-Bernoulli 1D raw-moment filter, TME-3 transition, N={N}, T={T}, '
                               f'bounded sample of {b} filters per step (full workload: {args.batch} per GPU)',
                   'N': N, 'T': T, 'batch_per_step': b, 'mode': 'raw', 'history': 'full'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': int(out['threads']), 'kind': 'port',
                         'sample': f'{b} filters x T={T} per step, {args.steps} steps; C restatement of the reference '
                                   f'algorithm (dense Cholesky + trsm + symmetric eigensolver), not JAX'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'diverged_frac': float(np.mean(out['status'] >= 0)),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='mfs_b200', choices=['mfs_b200', 'reference'])
    ap.add_argument('--N', type=int, default=8)
    ap.add_argument('--T', type=int, default=1000)
    ap.add_argument('--batch', type=int, default=1_000_000, help='filters per GPU')
    ap.add_argument('--mode', default='raw', choices=['raw', 'central'])
    ap.add_argument('--history', default='full', choices=['full', 'last', 'none'])
    ap.add_argument('--e2e-batch', type=int, default=131072, help='filters per GPU in the host-buffer (e2e) leg')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the theta-grid, 2-D filter and grid-filter legs')
    ap.add_argument('--secondary-scale', type=float, default=1.0,
                    help='record counts of the secondary legs relative to the stated BASELINE sizes (1/8 of them per rank)')
    ap.add_argument('--ref-step-seconds', type=float, default=10.)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import mfs_b200
    from mfs_b200 import _lib
    from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
    from mfs_b200.one_dim.moments import sde_cond_moments_tme
    from mfs_b200.one_dim.ss_models import benes_bernoulli
    from mfs_b200.simulate import simulate_1d

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)      # NCCL_DEBUG is left as the caller set it

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    N, T, B = args.N, args.T, args.batch
    M = 2 * N
    dt, _, _, ic, drift, disp, _, pmf, _ = benes_bernoulli(N)
    fam = sde_cond_moments_tme(drift, disp, dt, 3)

    # the seeded synthetic workload of this rank's shard (weak scaling: B filters per GPU), resident in HBM
    # (mfs_simulate_1d, exact Benes transition law; one Philox stream for the whole job: rank r owns trajectories
    # [r B, (r + 1) B), so the N-GPU job filters the same records whatever N is)
    ys = simulate_1d(drift, disp, dt, T, ic, pmf, B, 666 + 1, scheme='benes_exact', traj_offset=rank * B, device=dev)[2]
    out_bufs = {'nell': torch.empty(B, dtype=torch.float64, device=dev),
                'status': torch.empty(B, dtype=torch.int32, device=dev)}
    if args.history == 'full':
        out_bufs['ms'] = torch.empty((B, T, M), dtype=torch.float64, device=dev)
        if args.mode == 'central':
            out_bufs['mean'] = torch.empty((B, T), dtype=torch.float64, device=dev)
    elif args.history == 'last':
        out_bufs['ms'] = torch.empty((B, M), dtype=torch.float64, device=dev)
        if args.mode == 'central':
            out_bufs['mean'] = torch.empty((B,), dtype=torch.float64, device=dev)

    def step(ys_in, bufs, history, literal=False):
        if args.mode == 'raw':
            return moment_filter_rms(fam[0], pmf, ic.rms, ys_in, history=history, return_status=True, out=bufs,
                                     recompute_predict_quadrature=literal)
        return moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys_in, history=history, return_status=True,
                                 out=bufs, recompute_predict_quadrature=literal)

    def timed(fn, n):
        """n calls of fn bracketed by barriers; returns max-over-ranks milliseconds (CUDA events)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    fp64_peak, _ = mfs_b200.fp64_peak(local_rank, 4096)      # roofline denominator, measured on this GPU
    for _ in range(args.warmup):
        step(ys, out_bufs, args.history)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        res = step(ys, out_bufs, args.history)
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    status = res[-1]
    diverged = float((status >= 0).double().mean().item())
    live = float(torch.where(status >= 0, status, torch.full_like(status, T)).double().sum().item()) / (B * T)

    tmax = torch.tensor([total_ms, float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    agg = torch.tensor([diverged, live, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms, kernel_ms_avg = float(tmax[0]), float(tmax[1])
    diverged, live, launches_all = float(agg[0]) / world, float(agg[1]) / world, int(agg[2])
    steps_per_launch = B * T
    value = world * args.steps * steps_per_launch / (total_ms * 1e-3)

    # secondary figures (same buffers, same data): the literal two-quadrature recursion, and the reference's own
    # horizon T=100 (mfs/one_dim/ss_models.py:29), where ~1 % of the raw-moment filters diverge instead of ~43 %
    step(ys, out_bufs, args.history, literal=True)
    ms_lit = timed(lambda: step(ys, out_bufs, args.history, literal=True), args.steps)
    value_literal = world * args.steps * steps_per_launch / (ms_lit * 1e-3)
    kernel_ms_lit = ms_lit / args.steps
    T100 = min(100, T)
    ys100 = ys[:, :T100].contiguous()
    bufs100 = {'nell': out_bufs['nell'], 'status': out_bufs['status']}
    if args.history == 'full':
        bufs100['ms'] = out_bufs['ms'].view(-1)[:B * T100 * M]
        if args.mode == 'central':
            bufs100['mean'] = out_bufs['mean'].view(-1)[:B * T100]
    elif args.history == 'last':
        bufs100['ms'] = out_bufs['ms']
        if args.mode == 'central':
            bufs100['mean'] = out_bufs['mean']
    res100 = step(ys100, bufs100, args.history)
    ms100 = timed(lambda: step(ys100, bufs100, args.history), args.steps)
    value_t100 = world * args.steps * B * T100 / (ms100 * 1e-3)
    div100 = float((res100[-1] >= 0).double().mean().item())
    st100 = res100[-1]
    live100 = float(torch.where(st100 >= 0, st100, torch.full_like(st100, T100)).double().sum().item()) / (B * T100)

    # ---- e2e: host buffers through the public API (H2D + kernel + D2H every step) ----
    # Three output modes of the same call, weak-scaled (fixed filters per GPU):
    #   full     the reference's return value, (B, T, 2N) moments: 128 B per filter-step come back over PCIe
    #   meanvar  (mean, variance) per step -- what the reference's consumers read from the history: 16 B per step
    #   none     nell only (the estimation objective)
    e2e = None
    e2e_modes = {}
    if not args.no_e2e:
        def host_step(ys_np, bufs, history, chunk=0):
            if args.mode == 'raw':
                return moment_filter_rms(fam[0], pmf, ic.rms, ys_np, history=history, device=local_rank,
                                         return_status=True, out=bufs, chunk_filters=chunk)
            return moment_filter_cms(fam[1], fam[3], pmf, ic.cms, ic.mean, ys_np, history=history, device=local_rank,
                                     return_status=True, out=bufs, chunk_filters=chunk)

        def e2e_leg(history, Be, chunk=0):
            ys_host = torch.empty((Be, T), dtype=torch.uint8).pin_memory()
            ys_host.copy_(ys[:Be])
            bufs = {'nell': torch.empty(Be, dtype=torch.float64).pin_memory(),
                    'status': torch.empty(Be, dtype=torch.int32).pin_memory()}
            d2h = Be * 12
            if history == 'full':
                bufs['ms'] = torch.empty((Be, T, M), dtype=torch.float64).pin_memory()
                d2h += Be * T * M * 8
                if args.mode == 'central':
                    bufs['mean'] = torch.empty((Be, T), dtype=torch.float64).pin_memory()
                    d2h += Be * T * 8
            elif history == 'meanvar':
                bufs['meanvar'] = torch.empty((Be, T, 2), dtype=torch.float64).pin_memory()
                d2h += Be * T * 16
            ys_np = ys_host.numpy()
            host_step(ys_np, bufs, history, chunk)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                host_step(ys_np, bufs, history, chunk)      # returns after the last D2H completed
            barrier()
            el = time.perf_counter() - t0
            tt = torch.tensor([el], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            h2d = Be * T + 8 * (M + 8)
            gbs = (h2d + d2h) * args.steps / float(tt[0]) / 1e9
            return {'value': world * args.steps * Be * T / float(tt[0]), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'batch_per_gpu': Be, 'history': history,
                    'pcie_gb_per_s_per_gpu': gbs}

        prev_affinity, numa_cpus = bind_to_gpu_numa(local_rank)
        # full history: 16.8 GB of pinned host memory per 131072 filters and per rank
        Be_full = min(B, args.e2e_batch if world <= 2 else max(32768, args.e2e_batch // 4))
        e2e_modes['full'] = e2e_leg('full', Be_full)
        try:       # 8.4 GB of pinned host memory per rank; half of it if the host refuses
            e2e_modes['meanvar'] = e2e_leg('meanvar', min(B, 4 * args.e2e_batch), chunk=148 * 4 * 128)   # one full wave per chunk
        except RuntimeError:
            torch.cuda.empty_cache()
            e2e_modes['meanvar'] = e2e_leg('meanvar', min(B, 2 * args.e2e_batch), chunk=148 * 4 * 128)
        e2e_modes['none'] = e2e_leg('none', B)
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)
        e2e = dict(e2e_modes['full'])
        e2e['host_numa_binding'] = (f'rank bound to the {numa_cpus} CPUs NVML lists as local to its GPU while the pinned '
                                    f'buffers were allocated' if prev_affinity is not None else
                                    'none (NVML affinity unavailable, or all of this process\'s CPUs are local to the GPU)')
        e2e['note'] = ('pinned host buffers, chunked H2D->kernel->D2H pipeline inside mfs_filter_1d_host; the full moment '
                       'history (the reference\'s return value) is PCIe-bound at 128 B per filter-step; see e2e_modes for '
                       'the (mean, variance) history and the nell-only objective through the same call')
    e2e_nell = e2e_modes.get('none')

    secondary = None
    if not args.no_secondary:
        secondary = secondary_legs(rank, world, local_rank, fp64_peak, scale=args.secondary_scale)

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        W = work_per_step(N, two_quadratures=False)
        W_lit = work_per_step(N, two_quadratures=True)
        # flops actually executed: a filter that lost positive definiteness idles from its failing step on (NaN
        # outputs, like the JAX scan), so only LIVE filter-steps are claimed as work
        achieved_tf = W * steps_per_launch * live / (kernel_ms_avg * 1e-3) / 1e12
        alg_bytes = (1 + (8 * M if args.history == 'full' else 0)) * steps_per_launch
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                tj = json.load(f)
            key = f'N{N}_T{T}_B{B}_{args.mode}_{args.history}'
            traffic = tj.get(key)
        except Exception:
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'BASELINE configs[1]: Benes-Bernoulli 1D {args.mode}-moment filter, TME-3 '
                                   f'transition, N={N} nodes ({M} moments), T={T}, {B} independent filters per GPU',
                       'N': N, 'T': T, 'batch_per_gpu': B, 'global_batch': B * world, 'mode': args.mode,
                       'history': args.history, 'sharding': f'batch axis over {world} rank(s), no data-path collective',
                       'l2': 'inputs_larger_than_l2 (ys 1 GB, outputs 128 GB per step)', 'seed': 667,
                       'data_source': 'simulated on the device by mfs_simulate_1d (exact Benes transition law, '
                                      'Philox4x32-10 stream, trajectory ids sharded over ranks)'},
            'roofline': {'bound': 'fp64_fma', 'achieved': achieved_tf, 'peak': fp64_peak / 1e12, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / (fp64_peak / 1e12), 'traffic': traffic,
                         'traffic_note': 'dram__bytes_read + dram__bytes_write of the filter kernel(s) from one ncu --set full '
                                         'capture, per launch; see profiles/ncu_traffic.json for the batch it was captured at '
                                         '(scaled linearly to this batch when smaller) -- null when no capture matches',
                         'peak_source': 'measured live: mfs_fp64_peak DFMA micro-benchmark on this GPU '
                                        '(MEASURED_PEAKS.json has no FP64 figure; nominal 37.2)',
                         'flop_per_filter_step': W, 'kernel_ms': kernel_ms_avg, 'live_step_frac': live,
                         'achieved_counting_all_steps': W * steps_per_launch / (kernel_ms_avg * 1e-3) / 1e12,
                         'hbm': {'achieved': alg_bytes / (kernel_ms_avg * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'],
                                 'unit': 'GB/s', 'peak_source': peaks_src,
                                 'frac': alg_bytes / (kernel_ms_avg * 1e-3) / 1e9 / peaks['hbm_gbs'],
                                 'algorithmic_bytes_per_launch': alg_bytes}},
            'clocks': clocks, 'gpu_launches': launches_all, 'e2e': e2e, 'e2e_modes': e2e_modes, 'e2e_nell_only': e2e_nell,
            'diverged_frac': diverged, 'live_step_frac': live, 'value_live_steps_only': value * live,
            'wall_s_timed_region': t_wall,
            'literal_two_quadratures': {
                'value': value_literal, 'unit': UNIT, 'kernel_ms': kernel_ms_lit, 'flop_per_filter_step': W_lit,
                'roofline_frac': W_lit * steps_per_launch * live / (kernel_ms_lit * 1e-3) / fp64_peak,
                'note': 'MFS_FLAG_RECOMPUTE_PREDICT_QUADRATURE: second moment_quadrature per step, as filtering.py:78'},
            'reference_horizon_T100': {'value': value_t100, 'unit': UNIT, 'T': T100, 'diverged_frac': div100,
                                       'kernel_ms': ms100 / args.steps, 'live_step_frac': live100,
                                       'roofline_frac': W * B * T100 * live100 / (ms100 / args.steps * 1e-3) / fp64_peak,
                                       'note': 'the reference\'s own horizon (mfs/one_dim/ss_models.py:29): ~1 % of the '
                                               'filters diverge, so value and live-step value coincide'},
        }
        if secondary is not None:
            line['secondary'] = secondary
        if world == 1 and not args.no_cpu:
            v, cores, sample, _ = cpu_baseline(N, T)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': sample + '; C restatement of the reference algorithm, not JAX (scalar, one '
                                                       'filter per thread, no SIMD across filters)'}
            # BASELINE.md 3 "B2": the same dense algorithm as a batched torch-CPU fp64 program (MKL threads), an
            # independent cross-check of the C port's speed
            try:
                from baseline import torch_cpu_filter as B2
                from mfs_b200.synthetic import benes_bernoulli_ys_numpy
                bt, tt = 4096, min(T, 25)
                nell_t, el_t = B2.run(ic.rms[:M], benes_bernoulli_ys_numpy(bt, tt, 669))
                line['cpu_baseline_torch'] = {'value': bt * tt / el_t, 'unit': UNIT, 'cores': int(torch.get_num_threads()),
                                              'kind': 'port', 'sample': f'{bt} filters x T={tt}, batched torch.linalg '
                                              f'cholesky / solve_triangular / eigh on the host, {el_t:.1f} s'}
            except Exception as exc:            # a missing MKL path must not take the headline line down
                line['cpu_baseline_torch'] = {'unavailable': repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""BASELINE configs[3]: parameter-estimation objective over a theta grid (well--Poisson, central moments, TME-normal
order 2, N = 7 -- dardel/parameter_estimation/mf.py:21-53, run_parameter_estimation_mf.sh:34), argmin per trajectory.

Every (theta1, theta2, trajectory) triple is an independent filter; the flattened THETA grid is sharded over the ranks
(contiguous slices, no data-path collective), each rank reduces its slice to (min nell, argmin) per trajectory and the
pairs are combined with `mfs_b200.parallel.argmin_over_shards` (all-gather over NCCL; NCCL has no MINLOC).

    python tools/theta_grid_report.py [G] [n_traj] [T]                      # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/theta_grid_report.py ...
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mfs_b200 import synthetic
from mfs_b200.one_dim.filtering import moment_filter_cms
from mfs_b200.one_dim.moments import sde_cond_moments_tme_normal
from mfs_b200.one_dim.ss_models import well_poisson
from mfs_b200.parallel import shard_bounds, local_argmin, argmin_over_shards, gather_filters

G = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_traj = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
N = 7
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
dt, _, _, ic, drift, disp, emission, pmf, _ = well_poisson(3., N)
from mfs_b200.simulate import simulate_1d
ys_d = simulate_1d(drift(3.), disp, dt, T, ic, pmf(3.), n_traj, 670, device=dev)[2]      # simulated at theta = (3, 3); same on every rank
th1, th2 = np.meshgrid(np.linspace(0.5, 6., G), np.linspace(0.5, 6., G), indexing='ij')
th1, th2 = th1.reshape(-1), th2.reshape(-1)
lo, hi = shard_bounds(G * G, rank, world)
ys_b = ys_d[None].expand(hi - lo, n_traj, T)
fam = sde_cond_moments_tme_normal(drift(th1[lo:hi, None]), disp, dt, 2, N)
run = lambda: moment_filter_cms(fam[1], fam[3], pmf(th2[lo:hi, None]), ic.cms, ic.mean, ys_b, history='none',
                                return_status=True)
run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier(device_ids=[local])
torch.cuda.synchronize()
e0.record()
_, _, nell, status = run()
val, arg = local_argmin(nell, theta_offset=lo)
val, arg = argmin_over_shards(val, arg)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
# cross-check of the two collectives (outside the timed region): the all-gathered full nell table gives the same argmin
full = gather_filters(nell.reshape(hi - lo, n_traj), G * G)
val_f, arg_f = local_argmin(full)
assert full.shape == (G * G, n_traj) and torch.equal(arg_f, arg) and torch.equal(val_f, val), 'sharded argmin != argmin of the gathered table'
div = torch.tensor([float((status >= 0).sum()), float(status.numel())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(div)
if rank == 0:
    arg = arg.cpu().numpy()
    est = np.stack([th1[arg], th2[arg]], axis=1)
    steps = G * G * n_traj * T
    print(f'| grid | trajectories | T | GPUs | filters | filter-steps/s (incl. argmin) | diverged | median theta-hat | mean abs error of theta-hat |')
    print('|---|---|---|---|---|---|---|---|---|')
    print(f'| {G} x {G} over [0.5, 6]^2 | {n_traj} | {T} | {world} | {G * G * n_traj} | {steps / (float(ms[0]) * 1e-3):.3e} | '
          f'{float(div[0] / div[1]):.4f} | ({np.median(est[:, 0]):.2f}, {np.median(est[:, 1]):.2f}) | '
          f'({np.mean(np.abs(est[:, 0] - 3.)):.2f}, {np.mean(np.abs(est[:, 1] - 3.)):.2f}) |')
    print(f'\ncollectives ({"NCCL" if world > 1 else "single process"}, {world} rank(s)): argmin_over_shards == argmin of the gather_filters table: verified')
if world > 1:
    dist.destroy_process_group()

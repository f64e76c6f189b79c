"""world_size-2 gloo tests of the multi-GPU host logic (sharding, gathers, theta-grid argmin) on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mfs_b200 import parallel


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 8, 1000, 10 ** 6 + 3):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(total, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, n_theta, n_traj):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(total * 3, dtype=torch.float64).reshape(total, 3)
        mine = parallel.shard(full)
        lo, hi = parallel.shard_bounds(total, rank, world)
        assert mine.shape[0] == hi - lo
        gathered = parallel.gather_filters(mine * 2, total)
        assert torch.equal(gathered, full * 2)

        # theta grid sharded over ranks, argmin per trajectory; NaN (diverged) entries must never win
        g = torch.Generator().manual_seed(7)
        nell = torch.rand((n_theta, n_traj), generator=g, dtype=torch.float64)
        nell[3, :] = float('nan')
        nell[5, 2] = -1.          # unique minimum for trajectory 2
        tlo, thi = parallel.shard_bounds(n_theta, rank, world)
        lmin, larg = parallel.local_argmin(nell[tlo:thi], tlo)
        gmin, garg = parallel.argmin_over_shards(lmin, larg)
        clean = torch.where(torch.isnan(nell), torch.full_like(nell, float('inf')), nell)
        assert torch.equal(gmin, clean.min(dim=0).values)
        assert torch.equal(garg, clean.argmin(dim=0))
        assert garg[2].item() == 5
        # the same grid walked in chunks that do not fit one launch (bench.py secondary.theta_grid), then combined
        best = None
        for c0 in range(tlo, thi, 3):
            best = parallel.running_argmin(best, nell[c0:min(thi, c0 + 3)], c0)
        cmin, carg = parallel.argmin_over_shards(*best)
        assert torch.equal(cmin, gmin) and torch.equal(carg, garg)
        assert abs(parallel.max_over_ranks(float(rank)) - (world - 1)) < 1e-12
    finally:
        dist.destroy_process_group()


def test_gather_and_argmin_world2():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 11, 9, 4), nprocs=2, join=True)

"""BASELINE.md section 3, "B2": the reference's dense algorithm as a BATCHED torch-CPU fp64 program (MKL / OpenMP
threads) -- an independent cross-check of the speed of the C port (`oracle/mfs_oracle.c`, "B1") that bench.py reports as
`cpu_baseline`.  Benes--Bernoulli raw-moment filter only (the headline metric).

Per step and per filter, exactly the reference's operations (mfs/one_dim/quadtures.py:122-133, filtering.py:73-86):
Hankel gather, `torch.linalg.cholesky`, two `solve_triangular`, `torch.linalg.eigh` of the symmetrised K, weights =
V[0, :]^2, TME-3 transition moments (closed form for the Benes drift), Bernoulli likelihood,
renormalisation, nell.  Timing infrastructure only: bench.py calls `run`; nothing in mfs_b200/ imports this."""
import time

import numpy as np
import torch


def _quadrature(ms, n):
    idx = torch.arange(n)
    G = ms[:, idx[:, None] + idx[None, :]]
    H = ms[:, idx[:, None] + idx[None, :] + 1]
    R, info = torch.linalg.cholesky_ex(G)
    Y = torch.linalg.solve_triangular(R, H, upper=False)                       # R^-1 H
    K = torch.linalg.solve_triangular(R, Y.transpose(1, 2), upper=False).transpose(1, 2)
    K = 0.5 * (K + K.transpose(1, 2))
    lam, V = torch.linalg.eigh(K)
    return V[:, 0, :] ** 2, lam, info


def _benes_tme3(x, M, dt):
    """(B, n) nodes -> (B, n, M) transition moments E[X_dt^p | x]"""
    t = torch.tanh(x)
    dt2, dt3 = dt * dt, dt ** 3
    g = [None, dt * t, dt / 2 + dt2 / 2, (dt2 / 2 + dt3 / 6) * t, dt2 / 8 + dt3 / 4, dt3 / 8 * t, dt3 / 48]
    pw = [torch.ones_like(x)]
    for _ in range(1, M):
        pw.append(pw[-1] * x)
    cols = []
    for p in range(M):
        acc = pw[p]
        ff = 1.
        for k in range(1, min(p, 6) + 1):
            ff *= (p - k + 1)
            acc = acc + ff * g[k] * pw[p - k]
        cols.append(acc)
    return torch.stack(cols, dim=-1)


def run(rms0, ys, dt=1e-2, threads=None):
    """rms0 (2N,), ys (B, T) uint8 -> (nell (B,), seconds).  Failed filters (Cholesky info != 0) go NaN."""
    if threads:
        torch.set_num_threads(int(threads))
    ys = torch.as_tensor(np.asarray(ys))
    B, T = ys.shape
    M = len(rms0)
    n = M // 2
    ms = torch.as_tensor(np.asarray(rms0, dtype=np.float64)).repeat(B, 1)
    nell = torch.zeros(B, dtype=torch.float64)
    bad = torch.zeros(B, dtype=torch.bool)
    p = torch.arange(M, dtype=torch.float64)
    t0 = time.perf_counter()
    for t in range(T):
        w, x, info = _quadrature(ms, n)
        bad |= info != 0
        ms = torch.einsum('bi,bip->bp', w, _benes_tme3(x, M, dt))
        w, x, info = _quadrature(ms, n)
        bad |= info != 0
        pr = 1. / (1. + torch.exp(-x ** 3 / 5.))
        lik = torch.where(ys[:, t:t + 1] != 0, pr, 1. - pr)
        u = w * lik
        c = u.sum(dim=1, keepdim=True)
        ms = torch.einsum('bi,bip->bp', u, x.unsqueeze(-1) ** p) / c
        nell = nell - torch.log(c[:, 0])
        # keep failed filters from poisoning the batched LAPACK calls: park them on the initial moments
        ms = torch.where(bad[:, None], torch.as_tensor(np.asarray(rms0, dtype=np.float64)), ms)
    el = time.perf_counter() - t0
    nell[bad] = float('nan')
    return nell.numpy(), el

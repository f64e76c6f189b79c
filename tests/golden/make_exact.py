"""Exact-arithmetic fixture for the tolerance question (VERDICT r1 "weak #1"): the reference's Benes--Bernoulli scan
evaluated by the extended-precision arbiter ``oracle/mfs_oracle_mp.py`` (60 digits) on seeded records, rounded to
double.  fp64 implementations (LAPACK oracle, C oracle, CUDA kernel) are scored by their distance to these values.

    python tests/golden/make_exact.py            # ~10 minutes of mpmath; writes tests/golden/golden_exact_1d.npz

Inputs are fp64 values (dt = 0.01 as a double, the initial moments of mfs/one_dim/ss_models.py:25-35, uint8 records
from the exact Benes law), so every implementation sees bit-identical inputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import mfs_oracle as O        # noqa: E402
from oracle import mfs_oracle_mp as MP    # noqa: E402
from mfs_b200.synthetic import benes_bernoulli_ys_numpy  # noqa: E402

T, K = 100, 12


def main():
    out = {'T': T, 'K': K, 'digits': 60}
    for N in range(2, 9):
        ic = [r for r in O.benes_bernoulli(N) if hasattr(r, 'rms')][0]
        dt = 1e-2
        ys = benes_bernoulli_ys_numpy(K, T, 690 + N)
        out[f'N{N}/ys'] = ys
        out[f'N{N}/rms0'], out[f'N{N}/cms0'], out[f'N{N}/mean0'] = np.asarray(ic.rms[:2 * N]), np.asarray(ic.cms[:2 * N]), float(ic.mean)
        rm = np.full((K, T, 2 * N), np.nan)
        cm = np.full((K, T, 2 * N), np.nan)
        mn = np.full((K, T), np.nan)
        nr, nc = np.full(K, np.nan), np.full(K, np.nan)
        for k in range(K):
            h, nell = MP.moment_filter_rms(list(ic.rms[:2 * N]), ys[k], dt)
            for t, row in enumerate(h):
                if row is not None:
                    rm[k, t] = [float(v) for v in row]
            nr[k] = float(nell) if nell is not None else np.nan
            h, means, nell = MP.moment_filter_cms(list(ic.cms[:2 * N]), ic.mean, ys[k], dt)
            for t, row in enumerate(h):
                if row is not None:
                    cm[k, t] = [float(v) for v in row]
                    mn[k, t] = float(means[t])
            nc[k] = float(nell) if nell is not None else np.nan
            print(f'N={N} record {k}: nell raw {nr[k]:.15g} central {nc[k]:.15g}', flush=True)
        out[f'N{N}/rmss'], out[f'N{N}/nell_raw'] = rm, nr
        out[f'N{N}/cmss'], out[f'N{N}/means'], out[f'N{N}/nell_central'] = cm, mn, nc
    np.savez_compressed(os.path.join(HERE, 'golden_exact_1d.npz'), **out)
    print('golden_exact_1d.npz')


if __name__ == '__main__':
    main()

"""Graded-lexicographic multi-indices: enumeration, ranking and the Gram / Hankel gather tables of the d-dimensional
moment quadrature.  Host-side integer work done once per (N, d); results are bit-exact with
``mfs/multi_dims/multi_indices.py`` (pinned by ``tests/golden/golden_multi_indices.npz``).

Order (Dunkl & Xu 2014, p. 59): x > y iff |x| > |y|, or |x| = |y| and the first non-zero entry of x - y is positive.
E.g. d = 2: (0,0), (0,1), (1,0), (0,2), (1,1), (2,0), ...  -- position 1 is x_2, position 2 is x_1.
"""
import math
from typing import Sequence

import numpy as np

__all__ = ['sizeof_multi_indices', 'graded_lexico_indexof_multi_index', 'generate_graded_lexico_multi_indices',
           'find_indices', 'gram_and_hankel_indices_graded_lexico']


def _count_degree(d: int, s: int) -> int:
    """Number of d-dimensional multi-indices with |x| = s."""
    if d == 0:
        return 1 if s == 0 else 0
    return math.comb(s + d - 1, d - 1) if s >= 0 else 0


def sizeof_multi_indices(d: int, upper_sum: int, lower_sum: int = 0) -> int:
    """#{x in N^d : lower_sum <= |x| <= upper_sum}  (``multi_indices.py:25-58``)."""
    if upper_sum < lower_sum:
        return 0
    return sum(_count_degree(d, s) for s in range(lower_sum, upper_sum + 1))


def graded_lexico_indexof_multi_index(multi_index: Sequence[int], lower_sum: int = 0) -> int:
    """Position of ``multi_index`` in the graded-lex ordered collection {lower_sum <= |x|}  (``multi_indices.py:61-113``).

    Rank = (# indices of smaller degree) + (# indices of the same degree that are lexicographically smaller); the
    latter counts, for each position i and each value v < x_i, the completions of the remaining d-i-1 slots.
    """
    x = [int(v) for v in multi_index]
    d, s = len(x), sum(int(v) for v in multi_index)
    rank = sizeof_multi_indices(d, s - 1, lower_sum)
    remaining = s
    for i, xi in enumerate(x):
        for v in range(xi):
            rank += _count_degree(d - i - 1, remaining - v)
        remaining -= xi
    return rank


def generate_graded_lexico_multi_indices(d: int, upper_sum: int, lower_sum: int = 0) -> np.ndarray:
    """All multi-indices with lower_sum <= |x| <= upper_sum as an int64 (z, d) array  (``multi_indices.py:139-178``)."""
    rows = []

    def compositions(prefix, slots, total):
        if slots == 1:
            rows.append(prefix + [total])
            return
        for v in range(total + 1):              # ascending first entry = ascending graded-lex order
            compositions(prefix + [v], slots - 1, total - v)

    for s in range(lower_sum, upper_sum + 1):
        compositions([], d, s)
    return np.asarray(rows, dtype='int64').reshape(len(rows), d)


def find_indices(multi_indices) -> np.ndarray:
    """Vectorised rank (lower_sum = 0) over the leading axes of an (..., d) integer array  (``multi_indices.py:181-183``)."""
    arr = np.asarray(multi_indices)
    flat = arr.reshape(-1, arr.shape[-1])
    out = np.fromiter((graded_lexico_indexof_multi_index(r) for r in flat), dtype='int64', count=flat.shape[0])
    return out.reshape(arr.shape[:-1])


def gram_and_hankel_indices_graded_lexico(N: int, d: int) -> np.ndarray:
    """Gather tables (d + 1, s, s), s = C(N-1+d, N-1): ``ms[inds[0]]`` is the Gram matrix of the monomial basis of
    degree <= N-1, ``ms[inds[1 + i]]`` the multiplication-by-x_i (Hankel) matrix  (``multi_indices.py:185-229``)."""
    basis = generate_graded_lexico_multi_indices(d, N - 1, 0)
    pair = basis[:, None, :] + basis[None, :, :]
    s = basis.shape[0]
    inds = np.zeros((d + 1, s, s), dtype='int64')
    inds[0] = find_indices(pair)
    for i in range(d):
        shifted = pair.copy()
        shifted[..., i] += 1
        inds[i + 1] = find_indices(shifted)
    return inds

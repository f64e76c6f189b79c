"""The reference's per-run ``.npz`` result files, written from batched outputs.

The reference runs one Monte-Carlo record per process and stores one file per record:
``./results/benes_bernoulli_mf/{mode}{_t}_N_{N}_mc_{k}.npz`` with keys ``rmss, nell`` / ``cmss, means, nell`` /
``scmss, means, scales, nell`` (``dardel/benes_bernoulli/mf.py:83-92``) and
``./results/parameter_estimation_mf/N_{N}{_euler}_mc_{k}.npz`` with ``success, opt_params``
(``dardel/parameter_estimation/mf.py:76-77``).  The downstream scripts (``post_processing_mf.py``,
``compute_errs.py``, the plotting code) read exactly these names, so a batched run is split back into them here.
"""
import os

import numpy as np

__all__ = ['save_benes_bernoulli_mf', 'save_parameter_estimation_mf']

_KEYS = {'raw': ('rmss', 'nell'), 'central': ('cmss', 'means', 'nell'), 'scaled': ('scmss', 'means', 'scales', 'nell')}


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)


def save_benes_bernoulli_mf(results_dir, mode, N, outputs, transition=None, first_mc=0):
    """``outputs`` = the tuple a batched ``moment_filter_{rms,cms,scms}`` call returned (leading axis = Monte-Carlo
    record).  Writes one compressed file per record with the reference's file name and keys; ``transition`` is the
    reference's ``--transition`` suffix (``mf.py:34``: ``_t = f'_{transition}'`` or empty).  Returns the file names."""
    keys = _KEYS[mode]
    arrays = [_np(a) for a in outputs[:len(keys)]]
    _t = f'_{transition}' if transition else ''
    os.makedirs(results_dir, exist_ok=True)
    names = []
    for k in range(arrays[0].shape[0]):
        name = os.path.join(results_dir, f'{mode}{_t}_N_{N}_mc_{first_mc + k}.npz')
        np.savez_compressed(name, **{key: arr[k] for key, arr in zip(keys, arrays)})
        names.append(name)
    return names


def save_parameter_estimation_mf(results_dir, N, theta_hat, success, euler=False, first_mc=0):
    """One ``N_{N}{_euler}_mc_{k}.npz`` per run with ``success`` and ``opt_params`` (``mf.py:76-77``)."""
    theta_hat, success = _np(theta_hat), _np(success)
    os.makedirs(results_dir, exist_ok=True)
    names = []
    for k in range(theta_hat.shape[0]):
        name = os.path.join(results_dir, f'N_{N}{"_euler" if euler else ""}_mc_{first_mc + k}.npz')
        np.savez_compressed(name, success=bool(success[k]), opt_params=theta_hat[k])
        names.append(name)
    return names

"""mfs_b200 -- B200-native batched moment filter (drop-in for the hot path of zgbkdlm/mfs).

Layout mirrors the reference package for the functions on the path:
  mfs_b200.one_dim.filtering   moment_filter_rms / _cms / _scms     (mfs/one_dim/filtering.py)
  mfs_b200.one_dim.quadtures   moment_quadrature                    (mfs/one_dim/quadtures.py)
  mfs_b200.one_dim.moments     sde_cond_moments_* factories         (mfs/one_dim/moments.py)
  mfs_b200.one_dim.ss_models   benes_bernoulli, well_poisson        (mfs/one_dim/ss_models.py)
  mfs_b200.multi_dims          moment_filter_nd_rms / _cms, multi-index tables, nd factories, prey_predator
                                                                    (mfs/multi_dims/*)
  mfs_b200.classical_filters_smoothers.brute_force  brute_force_filter (mfs/classical_filters_smoothers/brute_force.py)
  mfs_b200.utils               GaussianSum1D                        (mfs/utils.py)
  mfs_b200.parallel            batch sharding + NCCL gathers        (new: the reference fans out OS processes)
The compute lives in ``mfs_b200/csrc`` (CUDA, sm_100a) behind the C ABI of ``include/mfs_b200.h``.
"""
from . import _lib
from ._lib import MfsError, build, launch_count, fp64_peak, dmma_peak
from . import functors

__version__ = '0.1.0'

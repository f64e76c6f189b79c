"""ctypes binding of ``libmfs_b200.so`` (C ABI in ``include/mfs_b200.h``).

There is no CPU fallback: if the shared library is missing or cannot be loaded every entry point raises.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MFS_B200_LIB') or os.path.join(_HERE, 'libmfs_b200.so')   # env override: tuning builds
CSRC = os.path.join(_HERE, 'csrc')

ABI_VERSION = 3
MAX_N = 15
MAX_PARAMS = 4

MODE = {'raw': 0, 'central': 1, 'scaled': 2}
TRANS = {'tme': 0, 'tme_normal': 1, 'euler': 2, 'normal_affine': 3}
DRIFT = {'benes': 0, 'well': 1, 'linear': 2}
MEAS = {'bernoulli_logistic_cubic': 0, 'poisson_softplus': 1, 'gaussian': 2}
YS_DTYPE = {'uint8': 0, 'bool': 0, 'int32': 1, 'float64': 2}
OUT_MODE = {'full': 0, 'last': 1, 'none': 2, 'meanvar': 3}
FLAG_RECOMPUTE_PREDICT_QUADRATURE = 1
GRAD_MAX_TANGENTS = 2


class Filter1dArgs(ctypes.Structure):
    """Mirror of ``mfs_filter1d_args``."""
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('mode', ctypes.c_int32), ('N', ctypes.c_int32), ('stable', ctypes.c_int32),
        ('B', ctypes.c_int64), ('T', ctypes.c_int64),
        ('trans_id', ctypes.c_int32), ('drift_id', ctypes.c_int32), ('tme_order', ctypes.c_int32),
        ('meas_id', ctypes.c_int32),
        ('dt', ctypes.c_double), ('dispersion', ctypes.c_double),
        ('trans_params', ctypes.c_void_p), ('trans_param_stride', ctypes.c_int64),
        ('meas_params', ctypes.c_void_p), ('meas_param_stride', ctypes.c_int64),
        ('ms0', ctypes.c_void_p), ('ms0_stride', ctypes.c_int64),
        ('mean0', ctypes.c_void_p), ('mean0_stride', ctypes.c_int64),
        ('scale0', ctypes.c_void_p), ('scale0_stride', ctypes.c_int64),
        ('ys', ctypes.c_void_p), ('ys_dtype', ctypes.c_int32), ('out_mode', ctypes.c_int32),
        ('ys_stride_b', ctypes.c_int64), ('ys_stride_t', ctypes.c_int64),
        ('ms_out', ctypes.c_void_p), ('ms_stride_b', ctypes.c_int64), ('ms_stride_t', ctypes.c_int64),
        ('mean_out', ctypes.c_void_p), ('scale_out', ctypes.c_void_p), ('aux_stride_b', ctypes.c_int64),
        ('nell_out', ctypes.c_void_p), ('status_out', ctypes.c_void_p),
        ('flags', ctypes.c_int32), ('segment_steps', ctypes.c_int32),
        ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_int64),
        ('carry_in', ctypes.c_void_p), ('carry_out', ctypes.c_void_p), ('t_offset', ctypes.c_int64),
        ('grid_records', ctypes.c_int64),
    ]


class FilterNdArgs(ctypes.Structure):
    """Mirror of ``mfs_filternd_args``."""
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('mode', ctypes.c_int32), ('N', ctypes.c_int32), ('d', ctypes.c_int32),
        ('B', ctypes.c_int64), ('T', ctypes.c_int64),
        ('trans_id', ctypes.c_int32), ('tme_order', ctypes.c_int32), ('meas_id', ctypes.c_int32),
        ('obs_dim', ctypes.c_int32), ('dt', ctypes.c_double),
        ('trans_params', ctypes.c_void_p), ('trans_param_stride', ctypes.c_int64),
        ('meas_params', ctypes.c_void_p), ('meas_param_stride', ctypes.c_int64),
        ('ms0', ctypes.c_void_p), ('ms0_stride', ctypes.c_int64),
        ('mean0', ctypes.c_void_p), ('mean0_stride', ctypes.c_int64),
        ('ys', ctypes.c_void_p), ('inds', ctypes.c_void_p),
        ('out_mode', ctypes.c_int32), ('stable', ctypes.c_int32),
        ('ms_out', ctypes.c_void_p), ('mean_out', ctypes.c_void_p), ('nell_out', ctypes.c_void_p),
        ('status_out', ctypes.c_void_p),
    ]


class BruteForceArgs(ctypes.Structure):
    """Mirror of ``mfs_brute_force_args``."""
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('pred_method', ctypes.c_int32), ('tme_order', ctypes.c_int32),
        ('integration_steps', ctypes.c_int32), ('n_grid', ctypes.c_int32), ('drift_id', ctypes.c_int32),
        ('meas_id', ctypes.c_int32), ('ys_dtype', ctypes.c_int32),
        ('B', ctypes.c_int64), ('T', ctypes.c_int64), ('dt', ctypes.c_double), ('dispersion', ctypes.c_double),
        ('trans_params', ctypes.c_void_p), ('meas_params', ctypes.c_void_p), ('meas_param_stride', ctypes.c_int64),
        ('xs', ctypes.c_void_p), ('init_ps', ctypes.c_void_p), ('init_ps_stride', ctypes.c_int64),
        ('ys', ctypes.c_void_p), ('ys_stride_b', ctypes.c_int64), ('ys_stride_t', ctypes.c_int64),
        ('out_mode', ctypes.c_int32), ('flags', ctypes.c_int32),
        ('pdfs_out', ctypes.c_void_p), ('nell_out', ctypes.c_void_p),
        ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_int64),
    ]


BF_METHOD = {'chapman-euler': 0, 'chapman-tme': 1, 'kolmogorov': 2}
BF_FLAG_POWER_OPERATOR = 1
SIM_SCHEME = {'tme': 0, 'benes_exact': 1}
SIM_MAX_COMPONENTS = 8


class Simulate1dArgs(ctypes.Structure):
    """Mirror of ``mfs_simulate1d_args``."""
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('scheme', ctypes.c_int32), ('tme_order', ctypes.c_int32),
        ('integration_steps', ctypes.c_int32), ('drift_id', ctypes.c_int32), ('meas_id', ctypes.c_int32),
        ('ys_dtype', ctypes.c_int32), ('n_components', ctypes.c_int32),
        ('B', ctypes.c_int64), ('T', ctypes.c_int64), ('dt', ctypes.c_double), ('dispersion', ctypes.c_double),
        ('trans_params', ctypes.c_void_p), ('trans_param_stride', ctypes.c_int64),
        ('meas_params', ctypes.c_void_p), ('meas_param_stride', ctypes.c_int64),
        ('init_means', ctypes.c_double * SIM_MAX_COMPONENTS), ('init_variances', ctypes.c_double * SIM_MAX_COMPONENTS),
        ('init_weights', ctypes.c_double * SIM_MAX_COMPONENTS),
        ('seed', ctypes.c_uint64), ('traj_offset', ctypes.c_uint64),
        ('ys_out', ctypes.c_void_p), ('ys_stride_b', ctypes.c_int64), ('ys_stride_t', ctypes.c_int64),
        ('xs_out', ctypes.c_void_p), ('xs_stride_b', ctypes.c_int64), ('xs_stride_t', ctypes.c_int64),
        ('x0_out', ctypes.c_void_p),
    ]


class SimulateLvArgs(ctypes.Structure):
    """Mirror of ``mfs_simulate_lv_args``."""
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('integration_steps', ctypes.c_int32), ('n_components', ctypes.c_int32),
        ('obs_dim', ctypes.c_int32), ('B', ctypes.c_int64), ('T', ctypes.c_int64), ('dt', ctypes.c_double),
        ('trans_params', ctypes.c_void_p), ('trans_param_stride', ctypes.c_int64),
        ('meas_params', ctypes.c_void_p), ('meas_param_stride', ctypes.c_int64),
        ('init_means', (ctypes.c_double * 2) * SIM_MAX_COMPONENTS), ('init_covs', (ctypes.c_double * 4) * SIM_MAX_COMPONENTS),
        ('init_weights', ctypes.c_double * SIM_MAX_COMPONENTS),
        ('seed', ctypes.c_uint64), ('traj_offset', ctypes.c_uint64),
        ('ys_out', ctypes.c_void_p), ('xs_out', ctypes.c_void_p), ('x0_out', ctypes.c_void_p),
    ]


EXPORTS = ('mfs_abi_version', 'mfs_last_error', 'mfs_functor_lookup', 'mfs_filter_1d', 'mfs_filter_1d_host',
           'mfs_moment_quadrature_1d', 'mfs_launch_count', 'mfs_fp64_peak', 'mfs_release_cached_memory', 'mfs_filter_nd',
           'mfs_brute_force', 'mfs_brute_force_workspace_bytes', 'mfs_brute_force_workspace_bytes_ex', 'mfs_dmma_peak', 'mfs_characteristic_fn_1d', 'mfs_filter_1d_workspace_bytes', 'mfs_moment_quadrature_nd',
           'mfs_simulate_1d', 'mfs_simulate_lv', 'mfs_filter_1d_grad', 'mfs_math_selftest')

_lib = None
_lock = threading.Lock()


class MfsError(RuntimeError):
    pass


def build(verbose: bool = False, jobs: int = None) -> str:
    """Compile the CUDA sources for sm_100a into ``mfs_b200/libmfs_b200.so`` (nvcc cross-compiles without a GPU)."""
    jobs = jobs or max(1, os.cpu_count() or 1)
    proc = subprocess.run(['make', '-C', CSRC, f'-j{jobs}'], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-8000:])
    if proc.returncode != 0:
        raise MfsError('building libmfs_b200.so failed (see output above)')
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises ``MfsError`` when it is absent -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MfsError(f'{LIB_PATH} not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                           f'(or `make -C mfs_b200/csrc -j`). There is no CPU fallback.')
        L = ctypes.CDLL(LIB_PATH)
        L.mfs_abi_version.restype = ctypes.c_int
        L.mfs_last_error.restype = ctypes.c_char_p
        L.mfs_functor_lookup.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32)]
        L.mfs_functor_lookup.restype = ctypes.c_int
        L.mfs_filter_1d.argtypes = [ctypes.POINTER(Filter1dArgs), ctypes.c_void_p]
        L.mfs_filter_1d.restype = ctypes.c_int
        L.mfs_filter_1d_grad.argtypes = [ctypes.POINTER(Filter1dArgs), ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                         ctypes.c_void_p, ctypes.c_void_p]
        L.mfs_filter_1d_grad.restype = ctypes.c_int
        L.mfs_filter_1d_workspace_bytes.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int64]
        L.mfs_filter_1d_workspace_bytes.restype = ctypes.c_int64
        L.mfs_filter_1d_host.argtypes = [ctypes.POINTER(Filter1dArgs), ctypes.c_int, ctypes.c_int64]
        L.mfs_filter_1d_host.restype = ctypes.c_int
        L.mfs_moment_quadrature_1d.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p]
        L.mfs_moment_quadrature_1d.restype = ctypes.c_int
        L.mfs_filter_nd.argtypes = [ctypes.POINTER(FilterNdArgs), ctypes.c_void_p]
        L.mfs_filter_nd.restype = ctypes.c_int
        L.mfs_brute_force.argtypes = [ctypes.POINTER(BruteForceArgs), ctypes.c_void_p]
        L.mfs_brute_force.restype = ctypes.c_int
        L.mfs_brute_force_workspace_bytes.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32]
        L.mfs_brute_force_workspace_bytes.restype = ctypes.c_int64
        L.mfs_brute_force_workspace_bytes_ex.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
        L.mfs_brute_force_workspace_bytes_ex.restype = ctypes.c_int64
        L.mfs_dmma_peak.argtypes = [ctypes.c_int, ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        L.mfs_dmma_peak.restype = ctypes.c_int
        L.mfs_characteristic_fn_1d.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 6
        L.mfs_characteristic_fn_1d.restype = ctypes.c_int
        L.mfs_moment_quadrature_nd.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.mfs_moment_quadrature_nd.restype = ctypes.c_int
        L.mfs_simulate_1d.argtypes = [ctypes.POINTER(Simulate1dArgs), ctypes.c_void_p]
        L.mfs_simulate_1d.restype = ctypes.c_int
        L.mfs_simulate_lv.argtypes = [ctypes.POINTER(SimulateLvArgs), ctypes.c_void_p]
        L.mfs_simulate_lv.restype = ctypes.c_int
        L.mfs_math_selftest.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 5
        L.mfs_math_selftest.restype = ctypes.c_int
        L.mfs_launch_count.restype = ctypes.c_int64
        L.mfs_fp64_peak.argtypes = [ctypes.c_int, ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        L.mfs_fp64_peak.restype = ctypes.c_int
        if L.mfs_abi_version() != ABI_VERSION:
            raise MfsError(f'ABI mismatch: library {L.mfs_abi_version()} vs binding {ABI_VERSION}')
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise MfsError(lib().mfs_last_error().decode() or f'libmfs_b200 returned {rc}')


def release_cached_memory():
    """Return the cached device staging buffers of the host-buffer path to the driver."""
    check(lib().mfs_release_cached_memory())


def launch_count() -> int:
    return int(lib().mfs_launch_count())


def fp64_peak(device: int = 0, iters: int = 4096):
    """Measured FP64 FMA throughput (FLOP/s) of ``device``; the roofline denominator of the filter kernels."""
    flops, ms = ctypes.c_double(), ctypes.c_double()
    check(lib().mfs_fp64_peak(device, iters, ctypes.byref(flops), ctypes.byref(ms)))
    return flops.value, ms.value


def dmma_peak(device: int = 0, iters: int = 4096):
    """Measured FP64 tensor-pipe (DMMA m8n8k4) throughput (FLOP/s) of ``device``: roofline denominator of the
    grid-filter GEMM."""
    flops, ms = ctypes.c_double(), ctypes.c_double()
    check(lib().mfs_dmma_peak(device, iters, ctypes.byref(flops), ctypes.byref(ms)))
    return flops.value, ms.value

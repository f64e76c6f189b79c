"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol of include/mfs_b200.h, functor
handles validate like the reference's API would, moment conversions match the reference-generated fixtures."""
import ctypes
import os
import re

import numpy as np
import pytest

import mfs_b200
from mfs_b200 import _lib, functors
from mfs_b200.one_dim import moments as M
from mfs_b200.one_dim.filtering import moment_filter_rms, moment_filter_cms
from mfs_b200.one_dim.ss_models import benes_bernoulli, well_poisson
from mfs_b200.utils import GaussianSum1D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'mfs_b200.h')).read()
    declared = set(re.findall(r'\b(mfs_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.mfs_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    # field order/count of the ctypes mirror vs the C struct in the header
    header = open(os.path.join(ROOT, 'include', 'mfs_b200.h')).read()
    body = header[header.index('typedef struct mfs_filter1d_args {'):header.index('} mfs_filter1d_args;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'[\w\*]+\s+\**(\w+);', body)
    assert names == [f[0] for f in _lib.Filter1dArgs._fields_]


def test_functor_lookup():
    L = _lib.lib()
    for kind, table in (('mode', _lib.MODE), ('trans', _lib.TRANS), ('drift', _lib.DRIFT), ('meas', _lib.MEAS)):
        for name, val in table.items():
            out = ctypes.c_int32(-7)
            assert L.mfs_functor_lookup(kind.encode(), name.encode(), ctypes.byref(out)) == 0
            assert out.value == val
    out = ctypes.c_int32()
    assert L.mfs_functor_lookup(b'drift', b'nope', ctypes.byref(out)) != 0
    assert b'unknown drift' in L.mfs_last_error()


def test_argument_validation_without_gpu():
    L = _lib.lib()
    a = _lib.Filter1dArgs()
    a.abi_version = 99
    assert L.mfs_filter_1d(ctypes.byref(a), None) != 0 and b'abi_version' in L.mfs_last_error()
    a.abi_version, a.N = _lib.ABI_VERSION, 1
    assert L.mfs_filter_1d(ctypes.byref(a), None) != 0 and b'N=1' in L.mfs_last_error()
    a.N, a.mode = 5, 7
    assert L.mfs_filter_1d(ctypes.byref(a), None) != 0 and b'mode' in L.mfs_last_error()
    a.mode, a.dt, a.tme_order, a.B, a.T = 0, 0.01, 3, 4, 3
    assert L.mfs_filter_1d(ctypes.byref(a), None) != 0 and b'NULL' in L.mfs_last_error()
    a.out_mode = 9
    assert L.mfs_filter_1d(ctypes.byref(a), None) != 0 and b'out_mode' in L.mfs_last_error()
    a.out_mode, a.t_offset = _lib.OUT_MODE['meanvar'], -1
    a.ms0 = a.trans_params = a.meas_params = a.nell_out = a.ys = a.ms_out = 8      # non-NULL, never dereferenced
    assert L.mfs_filter_1d(ctypes.byref(a), None) != 0 and b't_offset' in L.mfs_last_error()
    # the host-buffer entry point refuses the device-only carry before touching the GPU
    a.t_offset, a.carry_out = 0, 8
    assert L.mfs_filter_1d_host(ctypes.byref(a), 0, 0) != 0 and b'carry' in L.mfs_last_error()


def test_python_callables_are_rejected():
    dt, T, ts, ic, drift, disp, logistic, pmf, sim = benes_bernoulli(5)
    fam = M.sde_cond_moments_tme(drift, disp, dt, 3)
    ys = np.zeros(10, dtype=np.uint8)
    with pytest.raises(TypeError):
        moment_filter_rms(lambda x, n: x, pmf, ic.rms, ys)
    with pytest.raises(TypeError):
        moment_filter_rms(fam[0], lambda y, x: 1., ic.rms, ys)
    with pytest.raises(ValueError):
        moment_filter_rms(fam[1], pmf, ic.rms, ys)            # central member passed as raw
    other = M.sde_cond_moments_tme(drift, disp, dt, 3)
    with pytest.raises(ValueError):
        moment_filter_cms(fam[1], other[3], pmf, ic.cms, ic.mean, ys)   # members of different factory calls
    with pytest.raises(TypeError):
        M.sde_cond_moments_tme(np.tanh, disp, dt, 3)
    with pytest.raises(TypeError):
        fam[0](np.zeros(3), np.arange(4))
    with pytest.raises(ValueError):
        functors.Drift('unknown')


def test_pack_params():
    tab, stride = functors.pack_params((3., 2.), (4, 5))
    assert tab.shape == (1, 4) and stride == 0 and tab[0, 0] == 3. and tab[0, 1] == 2.
    th = np.arange(6.).reshape(2, 3)
    tab, stride = functors.pack_params((th, 7.), (2, 3))
    assert tab.shape == (6, 4) and stride == 4
    np.testing.assert_array_equal(tab[:, 0], th.reshape(-1))
    np.testing.assert_array_equal(tab[:, 1], 7.)


def test_conversions_and_initial_moments_match_reference_fixtures():
    g = np.load(os.path.join(GOLD, 'golden_conversions_1d.npz'))
    for N in (2, 5, 8):
        ic = GaussianSum1D.new([-0.5, 0.5], [0.05, 0.05], [0.5, 0.5], N)
        np.testing.assert_allclose(ic.rms, g[f'mix{N}/rms'], rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(ic.cms, g[f'mix{N}/cms'], rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(ic.scms, g[f'mix{N}/scms'], rtol=1e-13, atol=1e-16)
        assert abs(ic.mean - float(g[f'mix{N}/mean'])) < 1e-16 and abs(ic.variance - float(g[f'mix{N}/variance'])) < 1e-15
    rms = g['normal/rms']
    np.testing.assert_allclose(M.raw_to_central(rms), g['normal/raw_to_central'], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(M.raw_to_scaled(rms), g['normal/raw_to_scaled'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(M.central_to_raw(g['normal/raw_to_central'], 1.1), g['normal/central_to_raw'],
                               rtol=1e-12, atol=1e-9)
    # batched forms agree with the single-vector forms
    both = np.stack([rms, rms * 1.])
    np.testing.assert_array_equal(M.raw_to_central(both)[1], M.raw_to_central(rms))


def test_models_constants():
    dt, T, ts, ic, drift, disp, logistic, pmf, sim = benes_bernoulli(4)
    assert (dt, T) == (1e-2, 100) and drift.name == 'benes' and disp.value == 1. and pmf.params == (5., 0.)
    dt, T, ts, ic, drift, disp, emission, pmf, sim = well_poisson(3., 4)
    assert (dt, T) == (1e-2, 1000) and drift(2.5).params == (2.5,) and pmf(1.5).name == 'poisson_softplus'


def test_no_product_import_of_oracle():
    """The product package must never import, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'mfs_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', 'Makefile')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.lower().replace('no oracle', ''), os.path.join(dirpath, f)


def test_reference_result_files(tmp_path):
    """dardel/benes_bernoulli/mf.py:83-92 and dardel/parameter_estimation/mf.py:76-77: file names and keys."""
    from mfs_b200.results import save_benes_bernoulli_mf, save_parameter_estimation_mf
    rng = np.random.default_rng(0)
    cmss, means, nell = rng.normal(size=(3, 7, 4)), rng.normal(size=(3, 7)), rng.normal(size=3)
    names = save_benes_bernoulli_mf(tmp_path, 'central', 2, (cmss, means, nell), transition='tme_3', first_mc=10)
    assert [os.path.basename(n) for n in names] == [f'central_tme_3_N_2_mc_{k}.npz' for k in (10, 11, 12)]
    f = np.load(names[1])
    assert sorted(f.files) == ['cmss', 'means', 'nell']
    assert np.array_equal(f['cmss'], cmss[1]) and np.array_equal(f['means'], means[1]) and f['nell'] == nell[1]
    names = save_benes_bernoulli_mf(tmp_path, 'raw', 3, (cmss, nell))
    assert os.path.basename(names[0]) == 'raw_N_3_mc_0.npz' and sorted(np.load(names[0]).files) == ['nell', 'rmss']
    names = save_parameter_estimation_mf(tmp_path, 7, rng.normal(size=(2, 2)), np.array([True, False]), euler=True)
    f = np.load(names[1])
    assert os.path.basename(names[1]) == 'N_7_euler_mc_1.npz' and not bool(f['success']) and f['opt_params'].shape == (2,)

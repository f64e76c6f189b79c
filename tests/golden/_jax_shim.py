"""
A minimal NumPy-backed stand-in for the ``jax`` / ``tme`` modules, sufficient to EXECUTE the reference's own source
files (``/root/reference/mfs/**``) in a container that has no JAX.  Used only by ``make_golden.py`` to produce the
fixtures under ``tests/golden/`` (the reference tree does not travel to the GPU box, the fixtures do).

Semantics reproduced (float64 throughout, i.e. ``jax_enable_x64=True`` as every reference script sets):
  * ``jax.vmap(f, in_axes=...)``  - Python loop over the mapped axis, outputs stacked on axis 0 (pytree = tuple/list);
  * ``jax.lax.scan``              - Python loop, stacked outputs;
  * ``jax.lax.cond``              - evaluates the taken branch;
  * ``jax.lax.linalg.cholesky``   - LAPACK potrf, lower, NaN-filled result on failure (XLA semantics);
  * ``jax.lax.linalg.triangular_solve(a, b, left_side, lower, transpose_a)`` - LAPACK trtrs via SciPy;
  * ``jax.lax.linalg.eigh``       - symmetrises the input, LAPACK syevd, returns (vectors, values) like the jax
                                    vintage the reference targets (``mfs/one_dim/quadtures.py:131``);
  * ``jax.scipy.stats.{norm.pdf, norm.logpdf, bernoulli.pmf, poisson.pmf}`` - the formulas jax uses;
  * ``x.at[idx].set(v)``          - functional update on an ndarray subclass.
``tme.base_jax`` is installed as an empty placeholder: the third-party TME package is absent, so golden vectors never
go through it (transition moments come from the reference's Euler factory, or are passed in as callables).
"""
import sys
import types

import numpy as np
import scipy.linalg
import scipy.special


class ShimArray(np.ndarray):
    """ndarray with the functional ``.at[...]`` update API."""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = np.array(self.arr, copy=True).view(ShimArray)
        out[self.idx] = v
        return out

    def add(self, v):
        out = np.array(self.arr, copy=True).view(ShimArray)
        out[self.idx] += v
        return out


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, ShimArray):
        return x.view(ShimArray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    return x


class _NumpyProxy(types.ModuleType):
    """``jax.numpy``: forwards to numpy, returning ShimArray so that ``.at`` works."""

    def __getattr__(self, name):
        attr = getattr(np, name)
        if callable(attr) and not isinstance(attr, type):
            def fn(*args, **kwargs):
                with np.errstate(all='ignore'):
                    return _wrap(attr(*args, **kwargs))

            fn.__name__ = name
            return fn
        return attr


def _tree_stack(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([o[i] for o in outs]) for i in range(len(first)))
    return _wrap(np.stack([np.asarray(o) for o in outs], axis=0))


def _tree_index(tree, i):
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_index(t, i) for t in tree)
    return tree[i]


def _tree_len(tree):
    if isinstance(tree, (tuple, list)):
        return _tree_len(tree[0])
    return len(tree)


def vmap(f, in_axes=0, out_axes=0):
    assert out_axes == 0

    def wrapped(*args):
        axes = list(in_axes) if isinstance(in_axes, (list, tuple)) else [in_axes] * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            sliced = [a if ax is None else np.take(np.asarray(a), i, axis=ax) for a, ax in zip(args, axes)]
            outs.append(f(*sliced))
        return _tree_stack(outs)

    return wrapped


def scan(f, init, xs, length=None):
    carry = init
    n = _tree_len(xs) if xs is not None else length
    ys = []
    for i in range(n):
        carry, y = f(carry, _tree_index(xs, i) if xs is not None else None)
        ys.append(y)
    return carry, _tree_stack(ys)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def cholesky(x, symmetrize_input=True):
    x = np.asarray(x, dtype=np.float64)
    if symmetrize_input:
        x = 0.5 * (x + x.T)
    if not np.all(np.isfinite(x)):
        return _wrap(np.full_like(x, np.nan))
    try:
        return _wrap(scipy.linalg.cholesky(x, lower=True, check_finite=False))
    except scipy.linalg.LinAlgError:
        return _wrap(np.full_like(x, np.nan))


def triangular_solve(a, b, left_side=False, lower=False, transpose_a=False, conjugate_a=False, unit_diagonal=False):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not (np.all(np.isfinite(a)) and np.all(np.isfinite(b))):
        return _wrap(np.full_like(b, np.nan))
    trans = 1 if transpose_a else 0
    with np.errstate(all='ignore'):
        if left_side:       # solve op(a) x = b
            return _wrap(scipy.linalg.solve_triangular(a, b, trans=trans, lower=lower, unit_diagonal=unit_diagonal,
                                                       check_finite=False))
        # solve x op(a) = b  <=>  op(a)^T x^T = b^T
        return _wrap(scipy.linalg.solve_triangular(a, b.T, trans=1 - trans, lower=lower,
                                                   unit_diagonal=unit_diagonal, check_finite=False).T)


def eigh(x, lower=True, symmetrize_input=True, sort_eigenvalues=True):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 3:      # batched, like jax.lax.linalg.eigh
        pairs = [eigh(m, lower, symmetrize_input, sort_eigenvalues) for m in x]
        return _wrap(np.stack([p[0] for p in pairs])), _wrap(np.stack([p[1] for p in pairs]))
    if symmetrize_input:
        x = 0.5 * (x + x.T)
    if not np.all(np.isfinite(x)):
        n = x.shape[0]
        return _wrap(np.full((n, n), np.nan)), _wrap(np.full((n,), np.nan))
    w, v = scipy.linalg.eigh(x, lower=lower, driver='evd', check_finite=False)
    return _wrap(v), _wrap(w)


class _Norm:
    @staticmethod
    def logpdf(x, loc=0., scale=1.):
        z = (np.asarray(x, dtype=np.float64) - loc) / scale
        return _wrap(-0.5 * z * z - np.log(scale) - 0.5 * np.log(2 * np.pi))

    @staticmethod
    def pdf(x, loc=0., scale=1.):
        with np.errstate(all='ignore'):
            return _wrap(np.exp(_Norm.logpdf(x, loc, scale)))


class _Bernoulli:
    @staticmethod
    def logpmf(k, p, loc=0):
        with np.errstate(all='ignore'):
            k = np.asarray(k, dtype=np.float64) - loc
            return _wrap(scipy.special.xlogy(k, p) + scipy.special.xlog1py(1. - k, -np.asarray(p, dtype=np.float64)))

    @staticmethod
    def pmf(k, p, loc=0):
        with np.errstate(all='ignore'):
            return _wrap(np.exp(_Bernoulli.logpmf(k, p, loc)))


class _Poisson:
    @staticmethod
    def logpmf(k, mu, loc=0):
        with np.errstate(all='ignore'):
            k = np.asarray(k, dtype=np.float64) - loc
            return _wrap(scipy.special.xlogy(k, mu) - scipy.special.gammaln(k + 1.) - mu)

    @staticmethod
    def pmf(k, mu, loc=0):
        with np.errstate(all='ignore'):
            return _wrap(np.exp(_Poisson.logpmf(k, mu, loc)))


def _unavailable(name):
    def fn(*_, **__):
        raise NotImplementedError(f'{name} is not provided by the NumPy shim')

    return fn


def install():
    """Register the shim modules in ``sys.modules`` (idempotent)."""
    if 'jax' in sys.modules and getattr(sys.modules['jax'], '_is_numpy_shim', False):
        return
    jax = types.ModuleType('jax')
    jax._is_numpy_shim = True
    jnp = _NumpyProxy('jax.numpy')
    jnp.ndarray = np.ndarray
    jnp.linalg = np.linalg
    jax.numpy = jnp
    jax.Array = np.ndarray
    jax.vmap = vmap
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    jax.grad = _unavailable('jax.grad')
    jax.jacfwd = _unavailable('jax.jacfwd')
    jax.jacrev = _unavailable('jax.jacrev')
    jax.hessian = _unavailable('jax.hessian')

    lax = types.ModuleType('jax.lax')
    lax.scan, lax.cond = scan, cond
    lax.fori_loop = lambda lo, hi, body, init: _fori(lo, hi, body, init)
    lax_linalg = types.ModuleType('jax.lax.linalg')
    lax_linalg.cholesky, lax_linalg.triangular_solve, lax_linalg.eigh = cholesky, triangular_solve, eigh
    lax.linalg = lax_linalg
    jax.lax = lax

    jscipy = types.ModuleType('jax.scipy')
    jscipy_linalg = types.ModuleType('jax.scipy.linalg')
    jscipy_linalg.cholesky = lambda a, lower=False: cholesky(a) if lower else _wrap(np.asarray(cholesky(a)).T)
    jscipy_linalg.solve_triangular = lambda a, b, trans=0, lower=False: _wrap(
        scipy.linalg.solve_triangular(a, b, trans=trans, lower=lower, check_finite=False))
    jscipy_linalg.expm = lambda a: _wrap(scipy.linalg.expm(np.asarray(a)))
    jscipy_stats = types.ModuleType('jax.scipy.stats')
    jscipy_stats.norm, jscipy_stats.bernoulli, jscipy_stats.poisson = _Norm, _Bernoulli, _Poisson
    jscipy_special = types.ModuleType('jax.scipy.special')
    jscipy_special.gammaln = lambda x: _wrap(scipy.special.gammaln(x))
    jscipy.linalg, jscipy.stats, jscipy.special = jscipy_linalg, jscipy_stats, jscipy_special
    jax.scipy = jscipy

    config_mod = types.ModuleType('jax.config')

    class _Config:
        def update(self, *_):
            pass

    config_mod.config = _Config()
    jax.config = config_mod

    random = types.ModuleType('jax.random')
    for name in ('PRNGKey', 'split', 'normal', 'bernoulli', 'choice', 'uniform'):
        setattr(random, name, _unavailable(f'jax.random.{name}'))
    jax.random = random

    tme = types.ModuleType('tme')
    tme_base = types.ModuleType('tme.base_jax')
    tme_base.expectation = _unavailable('tme.expectation')
    tme_base.mean_and_cov = _unavailable('tme.mean_and_cov')
    tme.base_jax = tme_base

    for name, mod in [('jax', jax), ('jax.numpy', jnp), ('jax.lax', lax), ('jax.lax.linalg', lax_linalg),
                      ('jax.scipy', jscipy), ('jax.scipy.linalg', jscipy_linalg), ('jax.scipy.stats', jscipy_stats),
                      ('jax.scipy.special', jscipy_special), ('jax.config', config_mod), ('jax.random', random),
                      ('tme', tme), ('tme.base_jax', tme_base)]:
        sys.modules[name] = mod


def _fori(lo, hi, body, init):
    val = init
    for i in range(lo, hi):
        val = body(i, val)
    return val

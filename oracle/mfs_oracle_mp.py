"""Extended-precision arbiter for the 1-D moment-filter step (TEST INFRASTRUCTURE -- only tests/ and the fixture
generator tests/golden/make_exact.py import this; the product never does).

Why it exists.  The map moments -> quadrature is a Hankel Cholesky whose condition number grows like ~30^N, so two
faithful fp64 implementations of the reference's algorithm (LAPACK through NumPy, the C oracle, the CUDA kernel) differ
by cond * eps on unlucky trajectories, and "they agree to within their mutual noise" says nothing about WHO is closer
to the truth.  This module evaluates the reference's recursion in mpmath arithmetic (default 60 significant digits):
the result is, to double rounding, what the reference's algorithm returns in exact arithmetic on the same fp64 inputs.
Every fp64 implementation can then be scored by its distance to it.

Follows (file:line in /root/reference):
  mfs/one_dim/quadtures.py:122-133   Hankel gather, Cholesky, K = R^-1 H R^-T, symmetric eigen-decomposition,
                                     weights = V[0, :]^2, nodes = scale * lambda + mean
  mfs/one_dim/filtering.py:73-86     raw-moment scan body;  :140-158 central-moment scan body
  mfs/one_dim/ss_models.py:25-47     Benes drift tanh(x), b = 1, Bernoulli(1 / (1 + exp(-x^3 / 5)))
  mfs/one_dim/moments.py:141-179     TME transition moments; `tme` itself is absent (un-vendored, tme>=0.1.5), so the
                                     order-3 expansion is evaluated from its definition, here in the closed form the
                                     Benes identities a^2 + a' = 1, a a' + a''/2 = 0 give:
                                     T_p(x) = sum_{k<=6} c_{p,k}(tanh x) delta^{p-k},
                                     c_1 = dt p t, c_2 = p_2 (dt/2 + dt^2/2), c_3 = p_3 (dt^2/2 + dt^3/6) t,
                                     c_4 = p_4 (dt^2/8 + dt^3/4), c_5 = p_5 dt^3/8 t, c_6 = p_6 dt^3/48.
                                     tests/test_oracle_mp.py checks it against the sympy-derived TME of mfs_oracle.py.
"""
import mpmath as mp

__all__ = ['moment_quadrature', 'benes_tme_moments', 'moment_filter_rms', 'moment_filter_cms', 'set_precision']


def set_precision(digits: int = 60):
    mp.mp.dps = int(digits)


set_precision(60)


def moment_quadrature(ms, mean=0, scale=1):
    """(weights, nodes) as lists of mpf -- quadtures.py:122-133 in exact-like arithmetic."""
    n = len(ms) // 2
    G = mp.matrix(n, n)
    H = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            G[i, j] = ms[i + j]
            H[i, j] = ms[i + j + 1]
    R = mp.cholesky(G)                    # lower; raises ValueError when G is not positive definite
    Ri = mp.inverse(R)
    K = Ri * H * Ri.T
    K = (K + K.T) / 2                     # eigh(symmetrize_input=True)
    E, Q = mp.eigsy(K)
    return [Q[0, i] ** 2 for i in range(n)], [scale * E[i] + mean for i in range(n)]


def _falling(p, k):
    out = mp.mpf(1)
    for j in range(k):
        out *= (p - j)
    return out


def benes_tme_moments(x, num_moments, dt, order=3, m=0):
    """E[(X_dt - m)^p | x], p = 0..num_moments-1, TME of the given order for dX = tanh(X) dt + dW."""
    t = mp.tanh(x)
    d = x - m
    dt2, dt3 = dt * dt, dt * dt * dt
    o2, o3 = order >= 2, order >= 3
    g = [mp.mpf(1), dt * t,
         dt / 2 + (dt2 / 2 if o2 else 0),
         ((dt2 / 2 if o2 else 0) + (dt3 / 6 if o3 else 0)) * t,
         (dt2 / 8 if o2 else 0) + (dt3 / 4 if o3 else 0),
         (dt3 / 8 if o3 else 0) * t,
         (dt3 / 48 if o3 else 0)]
    out = []
    for p in range(num_moments):
        acc = mp.mpf(0)
        for k in range(min(p, 6) + 1):
            acc += _falling(p, k) * g[k] * d ** (p - k)
        out.append(acc)
    return out


def _bernoulli_logistic_cubic(y, x, c0=5, c1=0):
    p = 1 / (1 + mp.exp(-(x ** 3 / c0 - c1)))
    return p if y else 1 - p


def moment_filter_rms(rms0, ys, dt, order=3):
    """Raw-moment scan (filtering.py:73-86), Benes TME transition + Bernoulli likelihood.  Returns (rmss as a list of
    lists of mpf, nell); stops with None rows from the step whose Hankel matrix is not positive definite."""
    M = len(rms0)
    n = M // 2
    rms = [mp.mpf(v) for v in rms0]
    dt = mp.mpf(dt)
    nell = mp.mpf(0)
    hist = []
    for y in ys:
        try:
            w, x = moment_quadrature(rms)
            cols = [benes_tme_moments(xi, M, dt, order) for xi in x]
            rms = [sum(w[i] * cols[i][p] for i in range(n)) for p in range(M)]
            w, x = moment_quadrature(rms)
        except (ValueError, ZeroDivisionError):
            hist.extend([None] * (len(ys) - len(hist)))
            return hist, None
        lik = [_bernoulli_logistic_cubic(int(y), xi) for xi in x]
        c = sum(w[i] * lik[i] for i in range(n))
        rms = [sum(w[i] * x[i] ** p * lik[i] for i in range(n)) / c for p in range(M)]
        nell -= mp.log(c)
        hist.append(list(rms))
    return hist, nell


def moment_filter_cms(cms0, mean0, ys, dt, order=3):
    """Central-moment scan (filtering.py:140-158).  Returns (cmss, means, nell)."""
    M = len(cms0)
    n = M // 2
    cms = [mp.mpf(v) for v in cms0]
    mean = mp.mpf(mean0)
    dt = mp.mpf(dt)
    nell = mp.mpf(0)
    hist, means = [], []
    for y in ys:
        try:
            w, x = moment_quadrature(cms, mean)
            mean = sum(w[i] * (x[i] + dt * mp.tanh(x[i])) for i in range(n))      # TME mean (exact for Benes)
            cols = [benes_tme_moments(xi, M, dt, order, m=mean) for xi in x]
            cms = [sum(w[i] * cols[i][p] for i in range(n)) for p in range(M)]
            w, x = moment_quadrature(cms, mean)
        except (ValueError, ZeroDivisionError):
            hist.extend([None] * (len(ys) - len(hist)))
            means.extend([None] * (len(ys) - len(means)))
            return hist, means, None
        lik = [_bernoulli_logistic_cubic(int(y), xi) for xi in x]
        c = sum(w[i] * lik[i] for i in range(n))
        mean = sum(w[i] * x[i] * lik[i] for i in range(n)) / c
        cms = [sum(w[i] * (x[i] - mean) ** p * lik[i] for i in range(n)) / c for p in range(M)]
        nell -= mp.log(c)
        hist.append(list(cms))
        means.append(mean)
    return hist, means, nell

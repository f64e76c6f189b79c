"""Derive the TME operator coefficients of the Lotka--Volterra SDE symbolically (sympy) and print them as C.

For a smooth phi,  sum_{r<=M} dt^r/r! A^r phi = sum_{p,q} G_pq(x) d1^p d2^q phi,   A = a . grad + 1/2 Gamma : Hess,
a = (x1 (al - be x2), x2 (de x1 - ga)),  Gamma = diag(sg^2 x1^2, sg^2 x2^2)   (mfs/multi_dims/ss_models.py:55-59).
The printed G_pq are what mfs_b200/csrc/filter_nd.cuh: lv_tme_operator() evaluates; the oracle applies the generator to
every monomial separately instead, so the two derivations are independent.
usage: python tools/derive_lv_tme.py [order]
"""
import sys
import sympy as sp

order = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x1, x2, al, be, de, ga, sg2, dt = sp.symbols('x1 x2 al be de ga sg2 dt', real=True)
phi = sp.Function('phi')(x1, x2)
a1, a2 = x1 * (al - be * x2), x2 * (de * x1 - ga)
g11, g22 = sg2 * x1 ** 2, sg2 * x2 ** 2


def gen(f):
    return a1 * sp.diff(f, x1) + a2 * sp.diff(f, x2) + sp.Rational(1, 2) * (g11 * sp.diff(f, x1, 2) + g22 * sp.diff(f, x2, 2))


total, cur = phi, phi
for r in range(1, order + 1):
    cur = sp.expand(gen(cur))
    total = total + dt ** r / sp.factorial(r) * cur
total = sp.expand(total)
coeffs = {}
for p in range(2 * order + 1):
    for q in range(2 * order + 1 - p):
        if p == 0 and q == 0:
            d = phi
        else:
            args = ([(x1, p)] if p else []) + ([(x2, q)] if q else [])
            d = sp.Derivative(phi, *args)
        c = total.coeff(d)
        total = sp.expand(total - c * d)
        coeffs[(p, q)] = sp.factor(sp.simplify(c))
assert sp.simplify(total) == 0, total
names = sorted(coeffs)
repl, red = sp.cse([coeffs[k] for k in names], symbols=sp.numbered_symbols('t'))
for s, e in repl:
    print(f'const double {s} = {sp.ccode(e)};')
for k, e in zip(names, red):
    print(f'G[{k[0]}][{k[1]}] = {sp.ccode(e)};')

# degree-11 polynomial for exp(r) on [-ln2/2, ln2/2]: Chebyshev interpolation in high precision
import mpmath as mp
mp.mp.dps = 60
a = mp.log(2)/2 * mp.mpf('1.0001')
deg = 11
n = deg + 1
nodes = [a*mp.cos(mp.pi*(2*k+1)/(2*n)) for k in range(n)]
A = mp.matrix(n, n); b = mp.matrix(n, 1)
for i, x in enumerate(nodes):
    for j in range(n): A[i, j] = x**j
    b[i] = mp.e**x
c = mp.lu_solve(A, b)
# constrain c0 = 1, c1 = 1 exactly? check values
for j in range(n): print(j, mp.nstr(c[j], 20), float(c[j]).hex())
# error
import numpy as np
cf = [float(c[j]) for j in range(n)]
xs = np.linspace(-float(a), float(a), 20001)
p = np.zeros_like(xs)
for j in reversed(range(n)): p = p*xs + cf[j]
ref = np.array([float(mp.e**mp.mpf(x)) for x in xs[::50]])
print('max rel err (double eval)', np.max(np.abs(p[::50]-ref)/ref))
perr = max(abs(sum(c[j]*mp.mpf(x)**j for j in range(n)) - mp.e**mp.mpf(x))/mp.e**mp.mpf(x) for x in xs[::200])
print('approx error exact arith', mp.nstr(perr, 5))

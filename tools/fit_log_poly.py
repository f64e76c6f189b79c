# polynomial P(z), z = s^2 in [0, zmax], with atanh(s) = s (1 + z P(z)); s = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)]
import mpmath as mp
mp.mp.dps = 60
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1) * mp.mpf('1.001')
zmax = smax ** 2
deg = 7
n = deg + 1
nodes = [zmax / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]
def target(z):
    s = mp.sqrt(z)
    return (mp.atanh(s) / s - 1) / z
A = mp.matrix(n, n); b = mp.matrix(n, 1)
for i, z in enumerate(nodes):
    for j in range(n): A[i, j] = z ** j
    b[i] = target(z)
c = mp.lu_solve(A, b)
for j in range(n): print(j, mp.nstr(c[j], 20), float(c[j]).hex())
worst = 0
for k in range(1, 2000):
    z = zmax * k / 2000
    p = sum(c[j] * z ** j for j in range(n))
    s = mp.sqrt(z)
    approx = s * (1 + z * p)
    worst = max(worst, abs(approx - mp.atanh(s)) / mp.atanh(s))
print('rel err of atanh approx', mp.nstr(worst, 5))
print('ln2_hi', float(mp.log(2)).hex(), 'ln2_lo', float(mp.log(2) - mp.mpf(float(mp.log(2)))).hex())

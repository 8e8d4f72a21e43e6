"""Launch-bound regime: three bodies, nleap=100000 per output (reference tests/test_approx.py:6-17)."""
import sys, time
import numpy
sys.path.insert(0, '.')
import wendy_b200
x, v, m = numpy.array([-1.1, 0.1, 1.3]), numpy.array([3., 2., -5.]), numpy.array([1., 2., 3.])
for cap in (0, 2048):
    g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=100000, _cap=cap)
    next(g)
    t = time.perf_counter()
    for _ in range(3):
        tx, tv = next(g)
    el = time.perf_counter() - t
    print('cap=%d: %.2f us per sub-step' % (cap, el / 3e5 * 1e6), tx)
    g.close()
rs = numpy.random.RandomState(1)
S, L = 2000, 1000   # an ensemble of small systems: zero HBM traffic between sub-steps
x = rs.normal(size=S * L); v = rs.normal(size=S * L); m = numpy.full(S * L, 1. / L)
g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=100, n_segments=S)
next(g)
t = time.perf_counter(); next(g); el = time.perf_counter() - t
print('ensemble %d x %d: %.3e particle-steps/s' % (S, L, S * L * 100 / el))

"""GPU: the scenarios of the reference's own approximate-solver tests (tests/test_approx.py,
tests/test_approx_harm.py, tests/test_approx_external.py), re-written against wendy_b200 with
torch-vectorised external forces.  Invariants and thresholds are the reference's; only the number of
outputs / nleap is reduced where the reference spends 1e7 sub-steps on three particles."""
import numpy
import pytest

from oracle import wendy_oracle as wo

pytestmark = pytest.mark.gpu


def _disk(N=101, seed=2):
    rs = numpy.random.RandomState(seed)
    x = numpy.arctanh(2. * rs.uniform(size=N) - 1) * 2.
    v = rs.normal(size=N)
    v -= numpy.mean(v)
    m = numpy.ones(N) / N * (1. + 0.1 * (2. * rs.uniform(size=N) - 1))
    return x, v, m


def _conserves(gen, m, E, n_out, tol, **ekw):
    for _ in range(n_out):
        tx, tv = next(gen)
        assert abs(wo.energy(tx, tv, m, **ekw) - E) / abs(E) < tol
    gen.close()


@pytest.mark.parametrize('m', [[1., 1., 1.], [1., 2., 3.]])
def test_three_body_energy(m):  # tests/test_approx.py:6-32
    import wendy_b200
    x, v, m = numpy.array([-1.1, 0.1, 1.3]), numpy.array([3., 2., -5.]), numpy.array(m)
    # nleap as in the reference (the leapfrog error across collisions is first order in dt_leap); 3 outputs
    _conserves(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=100000), m, wo.energy(x, v, m), 3, 1e-6)


@pytest.mark.parametrize('sort', ['quick', 'merge', 'tim', 'qsort', 'parallel', 'gpu-radix'])
def test_disk_energy_for_every_reference_sort_name(sort):  # tests/test_approx.py:34-127
    import wendy_b200
    x, v, m = _disk()
    _conserves(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000, sort=sort), m, wo.energy(x, v, m), 20, 1e-6)


def test_disk_energy_harmonic():  # tests/test_approx_harm.py:6-53
    import wendy_b200
    x, v, m = _disk()
    E = wo.energy(x, v, m, omega=1.1)
    _conserves(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000, omega=1.1), m, E, 20, 1e-6, omega=1.1)


def test_external_force_as_harmonic_lambda_and_callable_class():  # tests/test_approx_external.py:5-63
    import wendy_b200
    x, v, m = _disk()
    omega = 1.1
    E = wo.energy(x, v, m, omega=omega)

    class Eforce(object):  # "a class, which numba can't handle" in the reference; any callable works here
        def __init__(self, omega):
            self._omega2 = omega ** 2.

        def __call__(self, x, t):
            return -self._omega2 * x
    for F in (lambda x, t: -omega ** 2. * x, Eforce(omega)):
        _conserves(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000, ext_force=F), m, E, 10, 1e-6, omega=omega)


def test_external_force_equals_builtin_harmonic():  # tests/test_approx_external.py:92-113 in spirit
    import wendy_b200
    x, v, m = _disk()
    g1 = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=200, omega=1.1)
    g2 = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=200, ext_force=lambda x, t: -1.1 ** 2. * x)
    for _ in range(5):
        a, b = next(g1), next(g2)
        assert numpy.max(numpy.abs(a[0] - b[0])) < 1e-12 and numpy.max(numpy.abs(a[1] - b[1])) < 1e-12
    g1.close(); g2.close()


def test_time_dependent_external_force_sees_the_reference_time_convention():
    """F(x,t)=c*t with no gravity (zero masses): v after a call is c*dt*sum_k (t0 + k dt)
    (force k of a call is evaluated at t0 + k*dt_leap, wendy/wendy.c:402-404)."""
    import wendy_b200
    n, c, t0, dt, nleap = 64, 0.3, 2.0, 0.1, 4
    x, v, m = numpy.linspace(-1, 1, n), numpy.zeros(n), numpy.zeros(n)
    g = wendy_b200.nbody(x, v, m, dt, approx=True, nleap=nleap, t0=t0, ext_force=lambda x, t: c * t + 0. * x)
    tx, tv = next(g)
    dtl = dt / nleap
    assert numpy.allclose(tv, c * dtl * sum(t0 + k * dtl for k in range(nleap)), rtol=1e-14, atol=0)
    tx, tv = next(g)  # t0 has advanced by dt
    assert numpy.allclose(tv, c * dtl * sum(t0 + k * dtl for k in range(2 * nleap)), rtol=1e-14, atol=0)
    g.close()


def test_momentum_conservation():  # tests/test_approx.py:151-163
    import wendy_b200
    x, v, m = _disk()
    v = v - numpy.sum(m * v) / numpy.sum(m)
    g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000)
    for _ in range(20):
        tx, tv = next(g)
        assert abs(wendy_b200.momentum(tv, m)) < 1e-10
    g.close()


def test_tracer_particles():  # tests/test_approx.py:165-185: zero-mass particles ride along
    import wendy_b200
    x, v, m = _disk()
    m[::5] = 0.
    _conserves(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000), m, wo.energy(x, v, m), 20, 1e-6)


def test_coincident_particles_at_the_first_force():  # tests/test_approx.py:234-248 (test_samex)
    import wendy_b200
    x, v, m = _disk()
    x[7], x[9] = x[3], x[3]
    _conserves(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000), m, wo.energy(x, v, m), 20, 1e-6)


def test_timer_and_missing_nleap():  # tests/test_approx.py:187-211
    import wendy_b200
    x, v, m = _disk()
    with pytest.raises(ValueError) as e:
        next(wendy_b200.nbody(x, v, m, 0.05, approx=True))
    assert str(e.value) == ('When approx is True, the number of leapfrog steps nleap= per output time step needs '
                            'to be set')
    g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=100, full_output=True)
    next(g)
    tx, tv, te = next(g)
    assert 0. < te < 1.
    g.close()

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the approximate-integration hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module, and only as the checker (or as the
timed CPU baseline).  The product path (``wendy_b200/``) never imports it.

Three checkers, strongest first:

``Reference``      the UNMODIFIED reference C path (``oracle/_ref/wendy_c.so``, compiled by
                   ``oracle/Makefile`` from the sources under /root/reference), driven through
                   the reference's own FFI entry point ``_wendy_nbody_approx_onestep``
                   (reference wendy/wendy.h:29-34) exactly as wendy/wendy.py:424-433 does --
                   except that the O(N) Python loop filling ``xi`` (wendy/wendy.py:384-387)
                   is replaced by a numpy structured array with the same 16-byte layout.
``COracle``        our C restatement (``oracle/wendy_oracle.c``).
``numpy_onestep``  a vectorised numpy restatement (np.cumsum is the same serial order as
                   wendy/wendy.c:359-360).

Parity status: PINNED -- see tests/test_oracle.py (golden vectors generated from
``Reference`` by tests/golden/make_golden.py, plus SURVEY.md 8(c) known answers).
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)
EXT_FORCE_CTYPE = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_int, _c_double_p,
                                   ctypes.c_double, _c_double_p)
#: numpy mirror of ``struct array_w_index`` (reference wendy/wendy.h:12-16): int32 idx,
#: 4 bytes padding, float64 val -> itemsize 16.
XI_DTYPE = numpy.dtype([('idx', 'i4'), ('val', 'f8')], align=True)
SORT_CODES = {'quick': 0, 'merge': 1, 'tim': 2, 'qsort': 3, 'parallel': 4}  # wendy/wendy.py:102


def _dp(arr):
    return arr.ctypes.data_as(_c_double_p)


def wrap_ext_force(ext_force):
    """ctypes adapter with the semantics of reference wendy/wendy.py:405-418."""
    if ext_force is None:
        return None

    def _cb(N, x, t, a):
        if N == 1:
            return float(ext_force(x.contents.value, t))
        xa = numpy.ctypeslib.as_array(x, shape=(N,))
        aa = numpy.ctypeslib.as_array(a, shape=(N,))
        aa[:] = ext_force(xa, t)
        return 0.
    return EXT_FORCE_CTYPE(_cb)


def reference_available(variant='wendy_c.so'):
    return os.path.exists(os.path.join(_HERE, '_ref', variant))


def c_oracle_available():
    return os.path.exists(os.path.join(_HERE, '_build', 'liboracle.so'))


class Reference(object):
    """Generator-style driver of the compiled reference (see module docstring)."""
    _libs = {}

    def __init__(self, x, v, m, dt, nleap, t0=0., twopiG=1., omega=None, ext_force=None,
                 sort='merge', variant='wendy_c.so'):
        if variant not in Reference._libs:
            lib = ctypes.CDLL(os.path.join(_HERE, '_ref', variant))
            lib._wendy_nbody_approx_onestep.restype = None
            lib._wendy_nbody_approx_onestep.argtypes = [
                ctypes.c_int, ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p,
                _c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, _c_double_p,
                ctypes.c_double, ctypes.c_void_p, ctypes.c_int, _c_int_p, _c_double_p,
                _c_double_p]
            Reference._libs[variant] = lib
        self._lib = Reference._libs[variant]
        # setup exactly as wendy/wendy.py:363-387,422
        self.omega2 = -1. if omega is None else omega ** 2.
        self.N = len(x)
        self.x = numpy.require(numpy.array(x, dtype='f8'), requirements=['C', 'W'])
        self.v = numpy.require(numpy.array(v, dtype='f8'), requirements=['C', 'W'])
        self.m = numpy.require(twopiG * numpy.array(m, dtype='f8'), requirements=['C', 'W'])
        self.a = numpy.zeros(self.N)
        self.cum = numpy.zeros(self.N)
        self.totmass = numpy.sum(self.m)
        self.xi = numpy.zeros(self.N, dtype=XI_DTYPE)
        self.xi['idx'] = numpy.arange(self.N, dtype='i4')
        self.xi['val'] = self.x
        self.t0 = ctypes.c_double(t0)
        self.err = ctypes.c_int(0)
        self.time_elapsed = ctypes.c_double(0.)
        self._cb = wrap_ext_force(ext_force)
        self.dt_leap = dt / nleap
        self.nleap = nleap
        self.sort = SORT_CODES[sort]

    def step(self):
        cb = ctypes.cast(self._cb, ctypes.c_void_p) if self._cb is not None else None
        self._lib._wendy_nbody_approx_onestep(
            self.N, self.xi.ctypes.data, _dp(self.x), _dp(self.v), _dp(self.m), _dp(self.a),
            self.totmass, self.dt_leap, self.nleap, ctypes.byref(self.t0), self.omega2,
            cb, self.sort, ctypes.byref(self.err), ctypes.byref(self.time_elapsed),
            _dp(self.cum))
        return self.x, self.v

    def __iter__(self):
        return self

    def __next__(self):
        return self.step()


class COracle(object):
    """Same driver, over our C restatement (oracle/wendy_oracle.c)."""
    _lib = None

    def __init__(self, x, v, m, dt, nleap, t0=0., twopiG=1., omega=None, ext_force=None):
        if COracle._lib is None:
            lib = ctypes.CDLL(os.path.join(_HERE, '_build', 'liboracle.so'))
            lib.oracle_approx_onestep.restype = None
            lib.oracle_approx_onestep.argtypes = [
                ctypes.c_int, _c_double_p, _c_int_p, _c_double_p, _c_double_p, _c_double_p,
                _c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, _c_double_p,
                ctypes.c_double, ctypes.c_void_p, _c_double_p]
            lib.oracle_argsort.restype = None
            lib.oracle_argsort.argtypes = [ctypes.c_int, _c_double_p, _c_int_p]
            COracle._lib = lib
        self.omega2 = -1. if omega is None else omega ** 2.
        self.N = len(x)
        self.x = numpy.array(x, dtype='f8')
        self.v = numpy.array(v, dtype='f8')
        self.m = twopiG * numpy.array(m, dtype='f8')
        self.a = numpy.zeros(self.N)
        self.cum = numpy.zeros(self.N)
        self.totmass = numpy.sum(self.m)
        self.sval = self.x.copy()
        self.sidx = numpy.arange(self.N, dtype='i4')
        self.t0 = ctypes.c_double(t0)
        self._cb = wrap_ext_force(ext_force)
        self.dt_leap = dt / nleap
        self.nleap = nleap

    def step(self):
        cb = ctypes.cast(self._cb, ctypes.c_void_p) if self._cb is not None else None
        COracle._lib.oracle_approx_onestep(
            self.N, _dp(self.sval), self.sidx.ctypes.data_as(_c_int_p), _dp(self.x),
            _dp(self.v), _dp(self.m), _dp(self.a), self.totmass, self.dt_leap, self.nleap,
            ctypes.byref(self.t0), self.omega2, cb, _dp(self.cum))
        return self.x, self.v

    def __iter__(self):
        return self

    def __next__(self):
        return self.step()


def argsort_key_then_index(x):
    """The framework's deterministic order: ascending x, ties by particle index.
    Coincides with every reference sort_type for distinct keys."""
    return numpy.lexsort((numpy.arange(len(x)), x))


def numpy_force(xs, ms, totmass, omega2, a_ext=None, cum=None):
    """Force on particles ALREADY in sorted order; reference wendy/wendy.c:359-383.

    ``cum`` overrides the serial running sum (used to test alternative scan orders)."""
    if cum is None:
        cum = numpy.concatenate(([0.], numpy.cumsum(ms)[:-1]))  # serial order == C loop
    g = (totmass - 2. * cum) - ms
    if omega2 >= 0:
        g = g - omega2 * xs
    return g if a_ext is None else a_ext + g


def numpy_onestep(x, v, m, totmass, dt, nleap, omega2=-1., ext_force=None, t0=0.,
                  exact_scan=False):
    """Vectorised restatement of reference wendy/wendy.c:385-418 on (x, v) given in
    particle-index order; returns (x, v, t0, last_sorted_ids).  ``m`` already includes
    twopiG.  ``exact_scan=True`` replaces the serial fp64 running sum by the correctly
    rounded exact prefix sum (what the CUDA path computes)."""
    x = numpy.array(x, dtype='f8')
    v = numpy.array(v, dtype='f8')
    x = x + (dt / 2.) * v
    order = None
    for k in range(nleap):
        order = argsort_key_then_index(x)
        xs, ms = x[order], m[order]
        a_ext = None if ext_force is None else numpy.asarray(ext_force(x, t0), dtype='f8')[order]
        cum = exact_prefix(ms) if exact_scan else None
        a = numpy_force(xs, ms, totmass, omega2, a_ext, cum)
        if ext_force is not None:
            t0 = t0 + dt
        vs = v[order] + dt * a
        xs = xs + (dt if k < nleap - 1 else dt / 2.) * vs
        v[order] = vs
        x[order] = xs
    return x, v, t0, order


def exact_prefix(ms):
    """Correctly rounded exclusive prefix sum (math.fsum semantics), O(N) via exact
    integer arithmetic on the fp64 significands."""
    if len(ms) == 0:
        return numpy.zeros(0)
    import math
    mant, expo = numpy.frexp(ms)
    emin = int(expo.min()) - 53
    ints = numpy.ldexp(mant, 53).astype(numpy.int64).tolist()  # exact 53-bit significands
    shifts = (expo - 53 - emin).tolist()
    out = numpy.empty(len(ms))
    acc = 0
    for i in range(len(ms)):
        # int -> float is correctly rounded (nearest even) in CPython; ldexp is then exact
        out[i] = math.ldexp(float(acc), emin)
        acc += ints[i] << shifts[i]
    return out


def energy(x, v, m, twopiG=1., omega=None):
    """System energy, reference wendy/wendy.py:458-475 (individual=False branch):
    harmonic + twopiG * sum_s m_s (M_below x_s - XM_below) + kinetic."""
    x = numpy.asarray(x, dtype='f8')
    v = numpy.asarray(v, dtype='f8')
    m = numpy.asarray(m, dtype='f8')
    harm = 0. if omega is None else numpy.sum(m * omega ** 2. * x ** 2. / 2.)
    s = numpy.argsort(x)
    below = numpy.concatenate(([0.], numpy.cumsum(m[s])[:-1]))
    xbelow = numpy.concatenate(([0.], numpy.cumsum((m * x)[s])[:-1]))
    return harm + twopiG * numpy.sum(m[s] * (below * x[s] - xbelow)) + numpy.sum(m * v ** 2. / 2.)


def potential(y, x, v, m, twopiG=1., omega=None, chunk=256):
    """reference wendy/wendy.py:494-517: omega^2 y^2/2 + twopiG * sum_i m_i |x_i - y_j| (O(N*Y) broadcast,
    evaluated in chunks of y so large N fits in memory; the row sums are numpy's pairwise sums as there)."""
    y = numpy.atleast_1d(numpy.asarray(y, dtype='f8'))
    x = numpy.asarray(x, dtype='f8')
    m = numpy.asarray(m, dtype='f8')
    out = numpy.empty(len(y))
    for s in range(0, len(y), chunk):
        yy = y[s:s + chunk]
        out[s:s + chunk] = twopiG * numpy.sum(m * numpy.fabs(x - numpy.atleast_2d(yy).T), axis=1)
    if omega is not None:
        out = omega ** 2. * y ** 2. / 2. + out
    return out


def energy_individual(x, v, m, twopiG=1., omega=None):
    """reference wendy/wendy.py:466-470 (individual=True branch)."""
    x = numpy.asarray(x, dtype='f8')
    v = numpy.asarray(v, dtype='f8')
    m = numpy.asarray(m, dtype='f8')
    out = 0. if omega is None else m * omega ** 2. * x ** 2. / 2.
    return out + m * potential(x, x, v, m, twopiG=twopiG) + m * v ** 2. / 2.


def momentum(v, m):
    """reference wendy/wendy.py:491"""
    return numpy.sum(numpy.asarray(m) * numpy.asarray(v))


# --- initial conditions used by the benchmark configs ------------
def sech2_ic(N, seed=2, zh=1., mass_jitter=0.):
    """Config 1/3: reference examples/WendyScaling.ipynb:57-65."""
    rs = numpy.random.RandomState(seed)
    x = numpy.arctanh(2. * rs.uniform(size=N) - 1.) * 2. * zh
    v = rs.normal(size=N)
    v -= numpy.mean(v)
    m = numpy.ones(N) / N
    if mass_jitter:
        m *= 1. + mass_jitter * (2. * rs.uniform(size=N) - 1.)
    return x, v, m


def slab_ic(N, seed=3):
    """Config 2 (survey-defined, the notebook is missing from the snapshot)."""
    rs = numpy.random.RandomState(seed)
    x = rs.uniform(-0.5, 0.5, size=N)
    v = 0.05 * rs.normal(size=N)
    v -= numpy.mean(v)
    return x, v, numpy.ones(N) / N

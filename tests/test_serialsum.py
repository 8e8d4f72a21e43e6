"""CPU: the closed form of the reference's SERIAL cumulative-mass sum (wendy_b200/csrc/serialsum.cuh, what the
equal-mass kernels evaluate) against plain serial summation -- numpy.cumsum (sequential in fp64) and the
cumulmass array the compiled reference itself leaves behind (/root/reference/wendy/wendy.c:359-360).
Host code only (wendy_serial_cum): no GPU needed."""
import math

import numpy
import pytest

from oracle import wendy_oracle as wo
from wendy_b200 import _lib


def _table(m0, n, k0=0):
    out = numpy.empty(n)
    pieces = _lib.load().wendy_serial_cum(float(m0), k0, n, out)
    assert pieces > 0, _lib.load().wendy_cuda_last_error()
    return out, pieces


def _serial(m0, n):
    ref = numpy.empty(n)
    ref[0] = 0.
    ref[1:] = numpy.cumsum(numpy.full(n - 1, m0))  # numpy's cumsum is the sequential fp64 recurrence
    return ref


def test_numpy_cumsum_is_the_serial_recurrence():
    m0, c = 0.1 / 3., 0.
    ref = _serial(m0, 5000)
    for i in range(5000):
        assert ref[i] == c
        c = c + m0


@pytest.mark.parametrize('m0', [1e-4, 2 * math.pi / 1e7, 0.3 / 1e5, 1. / 3., 3.0, 2. ** -20, 3 * 2. ** -20,
                                0.417022004702574, 7.2032449344215815e-06, 1. / 10000, 1. / 1000000])
def test_closed_form_equals_serial_sum(m0):
    n = 400000
    got, pieces = _table(m0, n)
    assert pieces < 200
    assert numpy.array_equal(got, _serial(m0, n))


def test_closed_form_random_masses_including_ties():
    rng = numpy.random.RandomState(5)
    n = 100000
    for t in range(150):
        m0 = float(rng.rand()) * 10. ** rng.randint(-12, 3)
        if t % 3 == 0:  # odd mantissa: exact round-half-even ties in the binade above
            mant, ex = math.frexp(m0)
            m0 = math.ldexp((int(mant * 2 ** 53) | 1) / 2. ** 53, ex)
        if t % 7 == 0:  # few significant bits: long runs without any rounding
            m0 = math.ldexp(float(rng.randint(1, 64)), int(rng.randint(-30, 5)))
        got, _ = _table(m0, n)
        assert numpy.array_equal(got, _serial(m0, n)), m0


def test_closed_form_at_the_headline_size_window():
    """N = 1e8 (BASELINE config 3): windows of the table far from the origin against the serial sum carried there."""
    m0 = 1. / 1e8
    n = 20000000
    ref = _serial(m0, n)
    got, _ = _table(m0, n)
    assert numpy.array_equal(got, ref)
    # continue the serial recurrence from position n-1 to check an offset window (k0 > 0) as the kernels use it
    tail = numpy.cumsum(numpy.concatenate([[ref[-1]], numpy.full(99999, m0)]))
    got2, _ = _table(m0, 100000, k0=n - 1)
    assert numpy.array_equal(got2, tail)


@pytest.mark.skipif(not wo.reference_available(), reason='oracle/_ref not built')
def test_closed_form_equals_the_compiled_reference_cumulmass():
    """The cumulmass scratch array the reference C code leaves behind after a step, equal masses with twopiG."""
    N = 200000
    x, v, m = wo.sech2_ic(N, seed=4)
    r = wo.Reference(x, v, m, 0.01, 1, twopiG=2. * math.pi, omega=0.9)
    r.step()
    got, _ = _table(r.m[0], N)
    assert numpy.array_equal(got, r.cum)

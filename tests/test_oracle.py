"""CPU: pin the oracle (C restatement + numpy restatement) against the golden vectors
generated from the unmodified reference (tests/golden/make_golden.py) and against the
known answers in SURVEY.md section 8(c)."""
import hashlib

import numpy
import pytest

from conftest import load_golden
from oracle import wendy_oracle as wo

CASES = ['kat_a', 'kat_b', 'kat_c', 'sech2_1000_nleap1', 'sech2_1000_nleap7_omega',
         'sech2_1000_twopiG', 'sech2_1000_ext', 'config1_sech2_1e4', 'slab_4096', 'ties',
         'tracers']
EXT = {'kat_c': lambda x, t: -1.21 * x + 0.1 * t,
       'sech2_1000_ext': lambda x, t: -0.7 * numpy.tanh(0.5 * x) + 0.05 * t}


def _kw(g, name):
    om = float(g['omega']) if 'omega' in g else numpy.nan
    return dict(omega=None if numpy.isnan(om) else om, ext_force=EXT.get(name),
                t0=float(g['t0']) if 't0' in g else 0.,
                twopiG=float(g['twopiG']) if 'twopiG' in g else 1.)


def _outputs(g):
    n = int(g['keep'][-1]) + 1 if 'keep' in g else len(g['xs'])
    keep = list(g['keep']) if 'keep' in g else list(range(n))
    return n, keep


@pytest.mark.parametrize('name', CASES)
def test_c_oracle_bit_exact_vs_reference_golden(name):
    g = load_golden(name)
    o = wo.COracle(g['x0'], g['v0'], g['m'], float(g['dt']), int(g['nleap']), **_kw(g, name))
    n, keep = _outputs(g)
    j = 0
    for i in range(n):
        x, v = o.step()
        if i in keep:
            assert numpy.array_equal(x, g['xs'][j]) and numpy.array_equal(v, g['vs'][j]), (name, i)
            j += 1
    if 'last_sorted_ids' in g:
        assert numpy.array_equal(o.sidx, g['last_sorted_ids'])


@pytest.mark.parametrize('name', [c for c in CASES if c != 'config1_sech2_1e4'])
def test_numpy_restatement_bit_exact_vs_reference_golden(name):
    g = load_golden(name)
    kw = _kw(g, name)
    m = kw['twopiG'] * g['m']
    omega2 = -1. if kw['omega'] is None else kw['omega'] ** 2.
    x, v, t0 = g['x0'], g['v0'], kw['t0']
    n, keep = _outputs(g)
    j = 0
    for i in range(n):
        x, v, t0, order = wo.numpy_onestep(x, v, m, numpy.sum(m), float(g['dt']) / int(g['nleap']),
                                           int(g['nleap']), omega2, kw['ext_force'], t0)
        if i in keep:
            assert numpy.array_equal(x, g['xs'][j]) and numpy.array_equal(v, g['vs'][j]), (name, i)
            j += 1


def test_exact_scan_variant_within_north_star_tolerance():
    """The CUDA path uses a correctly rounded exact prefix sum instead of the serial
    running sum; at these sizes the difference is far below the
    north-star tolerances (1e-12 after one step, 1e-9 after ten)."""
    g = load_golden('sech2_1000_nleap1')
    x, v = g['x0'], g['v0']
    for i in range(10):
        x, v, _, _ = wo.numpy_onestep(x, v, g['m'], numpy.sum(g['m']), 0.05, 1, exact_scan=True)
        tol = 1e-12 if i == 0 else 1e-9
        assert numpy.max(numpy.abs(x - g['xs'][i]) / numpy.maximum(1e-3, numpy.abs(g['xs'][i]))) < tol
        assert numpy.max(numpy.abs(v - g['vs'][i]) / numpy.maximum(1e-3, numpy.abs(g['vs'][i]))) < tol


def test_survey_known_answers():
    """hex literals for KAT-A/B/C and the KAT-D hashes."""
    fh = float.fromhex
    a = load_golden('kat_a')
    assert list(a['xs'][2]) == [fh('-0x1.2ffffffffffffp-1'), fh('0x1.b0a3d70a3d703p-2'), fh('0x1.0851eb851eb8cp-1')]
    assert list(a['vs'][2]) == [fh('0x1.dfffffffffffap+1'), fh('0x1.2666666666658p+1'), fh('-0x1.5ccccccccccc2p+2')]
    assert a['E'][0] == fh('0x1.f666666666667p+5') and a['E'][3] == fh('0x1.f66666666664cp+5')
    b = load_golden('kat_b')
    assert list(b['xs'][2]) == [fh('-0x1.2975ea3f1f29dp-1'), fh('0x1.adcde4b716fc3p-2'), fh('0x1.010f6678ddc35p-1')]
    c = load_golden('kat_c')
    assert list(c['vs'][2]) == [fh('0x1.f4e25a25a0227p+1'), fh('0x1.2187f926d6fa2p+1'), fh('-0x1.66cf058dfdb3bp+2')]
    d = load_golden('config1_sech2_1e4')
    assert hashlib.sha256(d['xs'][-1].tobytes()).hexdigest()[:16] == '21614814a2a156b6'
    assert hashlib.sha256(d['vs'][-1].tobytes()).hexdigest()[:16] == '7730018ad0084bb4'


def test_hand_derived_three_body_step():
    """SURVEY.md 8(c) sanity: equal masses, one output step dt=0.05 without crossings."""
    o = wo.COracle([-1.1, 0.1, 1.3], [3., 2., -5.], [1., 1., 1.], 0.05, 1)
    x, v = o.step()
    assert numpy.allclose(x, [-0.9475, 0.2, 1.0475], rtol=0, atol=1e-15)
    assert numpy.allclose(v, [3.1, 2., -5.1], rtol=0, atol=1e-15)


def test_energy_restatement_vs_reference_golden():
    for name in ['kat_a', 'kat_b', 'sech2_1000_nleap1', 'sech2_1000_nleap7_omega', 'sech2_1000_twopiG']:
        g = load_golden(name)
        kw = _kw(g, name)
        assert wo.energy(g['x0'], g['v0'], g['m'], kw['twopiG'], kw['omega']) == g['E'][0]
        assert wo.energy(g['xs'][-1], g['vs'][-1], g['m'], kw['twopiG'], kw['omega']) == g['E'][-1]


def test_argsort_ties_by_index():
    x = numpy.array([0.5, -1., 0.5, 0.5, -1., 2.])
    assert list(wo.argsort_key_then_index(x)) == [1, 4, 0, 2, 3, 5]
    sval, sidx = x.copy(), numpy.arange(6, dtype='i4')[::-1].copy()
    sval = x[sidx].copy()
    wo.COracle([0.], [0.], [1.], 1., 1)  # loads the library
    import ctypes
    wo.COracle._lib.oracle_argsort(6, sval.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                   sidx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert list(sidx) == [1, 4, 0, 2, 3, 5]


def test_exact_prefix_is_correctly_rounded():
    import math
    rs = numpy.random.RandomState(0)
    m = rs.uniform(size=200) * 10. ** rs.uniform(-8, 3, size=200)
    p = wo.exact_prefix(m)
    for i in (0, 1, 7, 100, 199):
        assert p[i] == math.fsum(m[:i])


@pytest.mark.parametrize('name', ['exact_solver_101', 'exact_solver_101_harm'])
def test_oracle_tracks_the_reference_exact_solver(name):
    """Physics cross-check (reference tests/test_approx.py:213-232, tests/test_approx_harm.py:55-75): the
    approximate integrator with nleap=2000 follows the reference's EXACT event-driven solver (golden
    fixture generated by tests/golden/make_golden.py) to 1e-5."""
    g = load_golden(name)
    om = None if numpy.isnan(float(g['omega'])) else float(g['omega'])
    o = wo.COracle(g['x0'], g['v0'], g['m'], 0.05, 2000, omega=om)
    for i in range(len(g['xs'])):
        x, v = o.step()
        assert numpy.max(numpy.abs(x - g['xs'][i])) < 1e-5 and numpy.max(numpy.abs(v - g['vs'][i])) < 1e-5


def test_potential_restatement_vs_reference_golden():
    """oracle potential / individual energies against wendy.potential / wendy.energy(individual=True)
    run in the dev container (tests/golden/make_golden.py)."""
    g = load_golden('potential_1000')
    tg, om = float(g['twopiG']), float(g['omega'])
    assert numpy.array_equal(wo.potential(g['y'], g['x'], g['v'], g['m'], twopiG=tg), g['pot'])
    assert numpy.array_equal(wo.potential(g['y'], g['x'], g['v'], g['m'], twopiG=tg, omega=om), g['pot_harm'])
    assert numpy.allclose(wo.energy_individual(g['x'], g['v'], g['m'], twopiG=tg), g['eind'], rtol=1e-14, atol=0)
    assert numpy.allclose(wo.energy_individual(g['x'], g['v'], g['m'], twopiG=tg, omega=om), g['eind_harm'],
                          rtol=1e-14, atol=0)

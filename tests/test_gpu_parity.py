"""GPU: parity of the CUDA path (through the C ABI, via wendy_b200.nbody / ctypes) against
  * the golden vectors generated from the unmodified reference (tests/golden), and
  * the CPU oracle on the same seeded inputs.

Tolerances (north star): sort permutation bit-exact; x, v relative 1e-12 after one step and
1e-9 after ten.  The relative error of an element is taken against max(|ref|, 0.1 rms(ref)) so
that a particle that happens to sit at x ~ 0 (or v ~ 0) does not turn an absolute 1e-15 into
a huge relative number.

What is actually required is stronger than those tolerances:
  * EQUAL masses (every BASELINE config): the CUDA path evaluates the reference's serial cumulative-mass
    sum in closed form, so x and v must be BIT-IDENTICAL to the reference C path -- checked against the
    goldens, the serial-scan restatement, and the compiled reference itself at N = 1e7 and 1e8
    (tests/test_gpu_reference_scale.py).
  * general masses: bit-identical to the oracle run with the SAME correctly rounded exact scan; the
    residual against the raw reference is the reference's own serial-summation error.
"""
import ctypes

import numpy
import pytest

from conftest import load_golden
from oracle import wendy_oracle as wo

pytestmark = pytest.mark.gpu

EXT_TORCH = {
    'kat_c': lambda x, t: -1.21 * x + 0.1 * t,
    'sech2_1000_ext': lambda x, t: -0.7 * __import__('torch').tanh(0.5 * x) + 0.05 * t,
}
SORTS = ['gpu', 'gpu-radix']
CAPS = [0, 2048]  # warp-per-bucket kernel (cap 256, default) and CTA-per-bucket kernel


def relerr(a, b):
    floor = 0.1 * numpy.sqrt(numpy.mean(b ** 2)) + 1e-300
    return numpy.max(numpy.abs(a - b) / numpy.maximum(floor, numpy.abs(b)))


def _kw(g, name):
    om = float(g['omega']) if 'omega' in g else numpy.nan
    return dict(omega=None if numpy.isnan(om) else om, ext_force=EXT_TORCH.get(name),
                t0=float(g['t0']) if 't0' in g else 0.,
                twopiG=float(g['twopiG']) if 'twopiG' in g else 1.)


# ---- sort --------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 33, 4096, 4097, 100003, 1 << 20])
def test_radix_argsort_bit_exact(n):
    import wendy_b200
    rs = numpy.random.RandomState(n)
    x = rs.normal(size=n) * 10. ** rs.randint(-3, 4, size=n)
    if n > 40:  # ties, signed zeros, extremes
        x[::7] = x[3]
        x[5], x[11] = 0., -0.
        x[13], x[17] = 1e308, -1e308
        x[19] = 5e-324
    perm = wendy_b200.argsort(x)
    assert numpy.array_equal(perm, wo.argsort_key_then_index(x))


def test_radix_argsort_matches_reference_first_sort():
    """The permutation the reference itself produced at its last force evaluation."""
    import wendy_b200
    g = load_golden('sech2_1000_nleap1')
    # one reference call with nleap=1: sorted at x0 + dt/2 v0
    x = g['x0'] + (0.05 / 2.) * g['v0']
    perm = wendy_b200.argsort(x)
    if wo.reference_available():
        r = wo.Reference(g['x0'], g['v0'], g['m'], 0.05, 1)
        r.step()
        assert numpy.array_equal(perm, r.xi['idx'])
    assert numpy.array_equal(perm, wo.argsort_key_then_index(x))


# ---- full path vs reference goldens --------------------------------------------------------
GOLDEN_CASES = ['kat_a', 'kat_b', 'kat_c', 'sech2_1000_nleap1', 'sech2_1000_nleap7_omega',
                'sech2_1000_twopiG', 'sech2_1000_ext', 'config1_sech2_1e4', 'slab_4096', 'ties',
                'tracers']


@pytest.mark.parametrize('cap', CAPS)
@pytest.mark.parametrize('sort', SORTS)
@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_against_reference_golden(name, sort, cap):
    import wendy_b200
    g = load_golden(name)
    n = int(g['keep'][-1]) + 1 if 'keep' in g else len(g['xs'])
    keep = list(g['keep']) if 'keep' in g else list(range(n))
    gen = wendy_b200.nbody(g['x0'], g['v0'], g['m'], float(g['dt']), approx=True,
                           nleap=int(g['nleap']), sort=sort, _cap=cap, **_kw(g, name))
    # equal masses and no transcendental in the external force: bit-identical to the reference's own output,
    # at every recorded output (up to 100 outputs x nleap sub-steps for config 1)
    bitwise = bool(numpy.all(g['m'] == g['m'][0])) and name != 'sech2_1000_ext'
    j = 0
    for i in range(n):
        x, v = next(gen)
        if i in keep:
            if bitwise:
                assert numpy.array_equal(x, g['xs'][j]) and numpy.array_equal(v, g['vs'][j]), (name, i)
            else:
                substeps = (i + 1) * int(g['nleap'])
                # north-star tolerances (general masses: exact scan vs the reference's serial sum)
                tol = 1e-12 if substeps <= 1 else 1e-9
                assert relerr(x, g['xs'][j]) < tol, (name, i, relerr(x, g['xs'][j]))
                assert relerr(v, g['vs'][j]) < tol, (name, i, relerr(v, g['vs'][j]))
            j += 1
    gen.close()


@pytest.mark.parametrize('sort', SORTS)
def test_first_steps_are_at_rounding_level(sort):
    """Well inside the tolerances: after one sub-step the GPU result differs from the reference
    only by the (better) rounding of the cumulative mass."""
    import wendy_b200
    g = load_golden('sech2_1000_nleap1')
    gen = wendy_b200.nbody(g['x0'], g['v0'], g['m'], 0.05, approx=True, nleap=1, sort=sort)
    x, v = next(gen)
    assert relerr(x, g['xs'][0]) < 1e-13 and relerr(v, g['vs'][0]) < 1e-13
    gen.close()


# ---- vs the oracle restatement with the SAME (exact) scan: bit-for-bit ----------------------
@pytest.mark.parametrize('cap', CAPS)
@pytest.mark.parametrize('sort', SORTS)
def test_bit_exact_vs_oracle_with_exact_scan(sort, cap):
    import wendy_b200
    x, v, m = wo.sech2_ic(3000, seed=7, mass_jitter=0.1)
    gen = wendy_b200.nbody(x, v, m, 0.02, approx=True, nleap=3, omega=0.9, sort=sort, _cap=cap)
    xo, vo = x, v
    for _ in range(3):
        xg, vg = next(gen)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.02 / 3, 3, 0.9 ** 2, exact_scan=True)
        assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)
    gen.close()


def test_bucket_and_radix_paths_agree_bitwise():
    import wendy_b200
    x, v, m = wo.sech2_ic(50000, seed=3, mass_jitter=0.2)
    outs = []
    for sort, cap in (('gpu', 0), ('gpu-radix', 0), ('gpu', 256)):
        gen = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=10, sort=sort, _cap=cap)
        for _ in range(3):
            xg, vg = next(gen)
        outs.append((xg.copy(), vg.copy()))
        gen.close()
    for xo, vo in outs[1:]:
        assert numpy.array_equal(outs[0][0], xo) and numpy.array_equal(outs[0][1], vo)


@pytest.mark.parametrize('n', [100000, 1000000])
def test_large_n_vs_oracle(n):
    """Config-2-like sizes, equal masses.  Default mode: bit-for-bit against the restatement of the reference
    (serial fp64 running sum, wendy.c:359-360) for ten sub-steps.  _exact_scan=True: bit-for-bit against the
    restatement with the correctly rounded exact scan; the two differ by the reference's own summation bias."""
    import wendy_b200
    x, v, m = wo.slab_ic(n, seed=3)
    gen = wendy_b200.nbody(x, v, m, 0.005, approx=True, nleap=1)
    gex = wendy_b200.nbody(x, v, m, 0.005, approx=True, nleap=1, _exact_scan=True)
    xo, vo = x, v
    xs, vs = x, v
    for i in range(10):
        xg, vg = next(gen)
        xe, ve = next(gex)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.005, 1, exact_scan=True)
        xs, vs, _, _ = wo.numpy_onestep(xs, vs, m, numpy.sum(m), 0.005, 1)
        assert numpy.array_equal(xg, xs) and numpy.array_equal(vg, vs), i
        assert numpy.array_equal(xe, xo) and numpy.array_equal(ve, vo), i
    bias = numpy.max(numpy.abs(numpy.cumsum(m) - numpy.arange(1, n + 1) / n))
    assert numpy.max(numpy.abs(ve - vs)) <= 10 * 0.005 * 2 * bias + 1e-15
    gen.close(); gex.close()


@pytest.mark.parametrize('sort', SORTS)
def test_equal_mass_specialisation_is_bit_identical_to_general_path(sort):
    """Equal masses with WENDY_FLAG_EXACT_SCAN take cum = RN(rank*m0); the general path takes the exact
    128-bit scan: the same correctly rounded numbers."""
    import wendy_b200
    x, v, m = wo.sech2_ic(30000, seed=8)
    outs = []
    for general in (False, True):
        gen = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=5, omega=1.1, sort=sort,
                               _general_masses=general, _exact_scan=not general)
        for _ in range(2):
            xg, vg = next(gen)
        outs.append((xg.copy(), vg.copy()))
        gen.close()
    assert numpy.array_equal(outs[0][0], outs[1][0]) and numpy.array_equal(outs[0][1], outs[1][1])


# ---- generator semantics ----------------------------------------------------------------------
def test_generator_yields_same_buffers_and_full_output():
    import wendy_b200
    x, v, m = wo.sech2_ic(500, seed=1)
    gen = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=2, full_output=True)
    a = next(gen)
    b = next(gen)
    assert a[0] is b[0] and a[1] is b[1]  # reference wendy/wendy.py:434-437
    assert 0. < b[2] < 1.  # reference tests/test_approx.py:198-211
    gen.close()


def test_nleap_call_pattern_matters_like_the_reference():
    """10 calls with nleap=1 differ from 1 call with nleap=10 at the 1e-14 level (SURVEY.md
    section 3.1); both must match the oracle run with the same pattern bit-for-bit."""
    import wendy_b200
    x, v, m = wo.sech2_ic(400, seed=9, mass_jitter=0.1)
    g1 = wendy_b200.nbody(x, v, m, 0.01, approx=True, nleap=1)
    for _ in range(10):
        x1, v1 = next(g1)
    g10 = wendy_b200.nbody(x, v, m, 0.1, approx=True, nleap=10)
    x10, v10 = next(g10)
    xo, vo = x, v
    for _ in range(10):
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.01, 1, exact_scan=True)
    assert numpy.array_equal(x1, xo) and numpy.array_equal(v1, vo)
    xo, vo, _, _ = wo.numpy_onestep(x, v, m, numpy.sum(m), 0.01, 10, exact_scan=True)
    assert numpy.array_equal(x10, xo) and numpy.array_equal(v10, vo)


@pytest.mark.parametrize('n', [1, 2, 127, 128, 129, 255, 256, 257, 1535, 1536, 1537, 2047, 2048, 2049, 4097])
def test_sizes_around_bucket_boundaries(n):
    import wendy_b200
    x, v, m = wo.sech2_ic(max(n, 2), seed=n, mass_jitter=0.1)
    x, v, m = x[:n], v[:n], m[:n]
    for cap in CAPS:
        gen = wendy_b200.nbody(x, v, m, 0.04, approx=True, nleap=2, omega=0.7, _cap=cap)
        xo, vo = x, v
        for _ in range(2):
            xg, vg = next(gen)
            xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.02, 2, 0.7 ** 2., exact_scan=True)
        gen.close()
        assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo), (n, cap)


def test_resume_from_a_yielded_state_is_bit_identical():
    """The reference's de-facto checkpoint/resume (examples/AdiabaticVsNonAdiabatic.ipynb:6791-6802): build a
    new generator from a yielded (x, v).  The yielded state is complete, and results do not depend on
    the internal bucket layout, so the continuation must be bit-identical."""
    import wendy_b200
    x, v, m = wo.sech2_ic(60000, seed=21, mass_jitter=0.05)
    g1 = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=5)
    for _ in range(3):
        xa, va = next(g1)
    g2 = wendy_b200.nbody(xa.copy(), va.copy(), m, 0.05, approx=True, nleap=5, sort='gpu-radix')
    for _ in range(2):
        xa, va = next(g1)
        xb, vb = next(g2)
    assert numpy.array_equal(xa, xb) and numpy.array_equal(va, vb)
    g1.close(); g2.close()


def test_interleaved_generators_are_independent():
    import wendy_b200
    ics = [wo.sech2_ic(5000, seed=s) for s in (1, 2)]
    gens = [wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=3) for x, v, m in ics]
    outs = [None, None]
    for _ in range(3):
        for i, g in enumerate(gens):
            outs[i] = [a.copy() for a in next(g)]
    for i, (x, v, m) in enumerate(ics):
        xo, vo = x, v
        for _ in range(3):
            xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.05 / 3, 3)  # equal masses: serial scan
        assert numpy.array_equal(outs[i][0], xo) and numpy.array_equal(outs[i][1], vo)
    [g.close() for g in gens]


def test_omega_zero_equals_no_harmonic_term():
    """reference wendy/wendy.py:363-366: omega=0.0 takes the harmonic branch with omega^2 = 0."""
    import wendy_b200
    x, v, m = wo.sech2_ic(3000, seed=3)
    a = next(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=4, omega=0.0))
    b = next(wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=4, omega=None))
    assert numpy.array_equal(a[0], b[0]) and numpy.array_equal(a[1], b[1])


# ---- energy / momentum -------------------------------------------------------------------------
@pytest.mark.parametrize('omega', [None, 1.1])
@pytest.mark.parametrize('n', [3, 1000, 100000])
def test_energy_matches_reference_formula(n, omega):
    import wendy_b200
    x, v, m = wo.sech2_ic(n, seed=4, mass_jitter=0.1)
    for tg in (1., 2.5):
        E = wo.energy(x, v, m, twopiG=tg, omega=omega)
        Eg = wendy_b200.energy(x, v, m, twopiG=tg, omega=omega)
        assert abs(Eg - E) <= 1e-12 * abs(E), (Eg, E)


def test_energy_and_momentum_conservation():
    """Reference tests/test_approx.py:34-51 (energy 1e-6 over 100 outputs), :151-163 (momentum)."""
    import wendy_b200
    N = 101
    rs = numpy.random.RandomState(2)
    x = numpy.arctanh(2. * rs.uniform(size=N) - 1) * 2.
    v = rs.normal(size=N)
    v -= numpy.mean(v)
    m = numpy.ones(N) / N * (1. + 0.1 * (2. * rs.uniform(size=N) - 1))
    v -= numpy.sum(m * v) / numpy.sum(m)
    E = wendy_b200.energy(x, v, m)
    gen = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1000)
    for _ in range(20):
        tx, tv = next(gen)
        assert abs(wo.energy(tx, tv, m) - E) / abs(E) < 1e-6
        assert abs(wendy_b200.momentum(tv, m)) < 1e-10
    gen.close()


# ---- potential(y) and per-particle energies (reference wendy/wendy.py:494-517, 466-470) ----------------
def test_potential_and_individual_energies_vs_reference_golden():
    """fp64 tolerance: 1e-13 of the largest potential (different, fixed summation order)."""
    import wendy_b200
    g = load_golden('potential_1000')
    tg, om = float(g['twopiG']), float(g['omega'])
    for omega, kp, ke in ((None, 'pot', 'eind'), (om, 'pot_harm', 'eind_harm')):
        p = wendy_b200.potential(g['y'], g['x'], g['v'], g['m'], twopiG=tg, omega=omega)
        assert numpy.max(numpy.abs(p - g[kp])) <= 1e-13 * numpy.max(numpy.abs(g[kp]))
        e = wendy_b200.energy(g['x'], g['v'], g['m'], twopiG=tg, individual=True, omega=omega)
        assert numpy.max(numpy.abs(e - g[ke])) <= 1e-13 * numpy.max(numpy.abs(g[ke]))


@pytest.mark.parametrize('n', [1, 2, 2047, 2048, 2049, 300000])
def test_potential_vs_oracle_sizes(n):
    import wendy_b200
    x, v, m = wo.sech2_ic(n, seed=11, mass_jitter=0.3)
    y = numpy.concatenate((numpy.linspace(-9., 9., 57), x[:: max(1, n // 13)], [-1e3, 1e3]))
    ref = wo.potential(y, x, v, m, twopiG=2., omega=0.7)
    got = wendy_b200.potential(y, x, v, m, twopiG=2., omega=0.7)
    assert numpy.max(numpy.abs(got - ref)) <= 1e-12 * numpy.max(numpy.abs(ref))
    assert wendy_b200.potential(numpy.empty(0), x, v, m).shape == (0,)


def test_individual_energies_sum_rule_large_n_device_tensors():
    """sum_i E_i = kinetic + harmonic + 2 * pair energy, so sum_i E_i - energy() is the pair energy; torch
    CUDA tensors in, CUDA tensor out (nothing staged through the host)."""
    import torch
    import wendy_b200
    n = 2000000
    x, v, m = wo.sech2_ic(n, seed=5, mass_jitter=0.2)
    xd, vd, md = (torch.as_tensor(a, device='cuda') for a in (x, v, m))
    e = wendy_b200.energy(xd, vd, md, twopiG=1.3, individual=True, omega=0.9)
    assert e.is_cuda and e.shape == (n,)
    E = wo.energy(x, v, m, twopiG=1.3, omega=0.9)
    one_body = numpy.sum(m * v ** 2. / 2.) + numpy.sum(m * 0.81 * x ** 2. / 2.)
    pair = E - one_body
    assert abs(float(e.sum()) - (one_body + 2. * pair)) <= 1e-11 * abs(E)
    yd = torch.linspace(-5., 5., 1001, dtype=torch.float64, device='cuda')
    p = wendy_b200.potential(yd, xd, vd, md, twopiG=1.3)
    assert p.is_cuda
    ref = wo.potential(yd.cpu().numpy()[::100], x, v, m, twopiG=1.3)
    assert numpy.max(numpy.abs(p.cpu().numpy()[::100] - ref)) <= 1e-12 * numpy.max(numpy.abs(ref))
    # the potential is convex with slope -> +-twopiG*M outside the system
    pp = p.cpu().numpy()
    assert numpy.all(numpy.diff(pp, 2) >= -1e-12)


# ---- physics cross-check against the reference's exact solver ------------------------------------
@pytest.mark.parametrize('name', ['exact_solver_101', 'exact_solver_101_harm'])
def test_tracks_the_reference_exact_solver(name):
    """reference tests/test_approx.py:213-232 / tests/test_approx_harm.py:55-75, against a golden run of
    the reference's event-driven solver (out of scope here, used only as a physics cross-check)."""
    import wendy_b200
    g = load_golden(name)
    om = None if numpy.isnan(float(g['omega'])) else float(g['omega'])
    gen = wendy_b200.nbody(g['x0'], g['v0'], g['m'], 0.05, approx=True, nleap=2000, omega=om)
    for i in range(len(g['xs'])):
        x, v = next(gen)
        assert numpy.max(numpy.abs(x - g['xs'][i])) < 1e-5 and numpy.max(numpy.abs(v - g['vs'][i])) < 1e-5
    gen.close()


# ---- ensembles of independent realisations -------------------------------------------------------
@pytest.mark.parametrize('sort', SORTS)
def test_segments_equal_independent_runs(sort):
    import wendy_b200
    S, L = 5, 700
    xs, vs, ms = zip(*[wo.sech2_ic(L, seed=20 + s, mass_jitter=0.1) for s in range(S)])
    gen = wendy_b200.nbody(numpy.concatenate(xs), numpy.concatenate(vs), numpy.concatenate(ms),
                           0.05, approx=True, nleap=4, omega=0.5, sort=sort, n_segments=S, _cap=256)
    for _ in range(3):
        X, V = next(gen)
    gen.close()
    for s in range(S):
        g1 = wendy_b200.nbody(xs[s], vs[s], ms[s], 0.05, approx=True, nleap=4, omega=0.5, sort=sort)
        for _ in range(3):
            x1, v1 = next(g1)
        g1.close()
        assert numpy.array_equal(X[s * L:(s + 1) * L], x1) and numpy.array_equal(V[s * L:(s + 1) * L], v1)


def test_large_ensemble_on_the_persistent_kernel_equals_the_oracle():
    """Segments on 2048-slot buckets with MORE buckets than resident CTAs: a persistent CTA then walks buckets of
    several segments one after the other, and everything it carries between buckets (the piece of the serial
    cumulative-mass table, the splitter window, the prefetched counts) must follow the segment change.  Equal
    masses: bit-identical to the C restatement of the reference, per segment; no sub-step may fail over."""
    import wendy_b200
    S, L = 4, 300000  # 1.2e6 particles: the library chooses the coarse geometry; 4 x 181 buckets > 2 x 148 CTAs
    ics = [wo.sech2_ic(L, seed=60 + s) for s in range(S)]
    X = numpy.concatenate([i[0] for i in ics]); V = numpy.concatenate([i[1] for i in ics])
    M = numpy.concatenate([i[2] for i in ics])
    st = wendy_b200.ApproxState(X, V, M, omega2=1.1 ** 2., n_segments=S)
    for _ in range(2):
        st.step(0.004, 3)
    Xg, Vg = st.read()
    stats = st.stats()
    st.close()
    assert stats['cap'] == 2048 and stats['buckets'] > 296
    assert stats['failed_substeps'] == 0 and stats['radix_fallbacks'] == 0, stats
    for s in range(S):
        ref = wo.COracle(ics[s][0], ics[s][1], ics[s][2], 0.012, 3, omega=1.1)
        for _ in range(2):
            xr, vr = ref.step()
        assert numpy.array_equal(Xg[s * L:(s + 1) * L], xr) and numpy.array_equal(Vg[s * L:(s + 1) * L], vr), s


# ---- compat export: the reference's own C entry point --------------------------------------------
def _compat_call(lib, xi, x, v, m, tot, dt, nleap, t0, omega2, cb):
    from wendy_b200 import _lib
    a = numpy.zeros(len(x))
    cum = numpy.zeros(len(x))
    err = ctypes.c_int(0)
    te = ctypes.c_double(0.)
    t0c = ctypes.c_double(t0)
    lib._wendy_nbody_approx_onestep(len(x), xi.ctypes.data, x, v, m, a, tot, dt, nleap,
                                    ctypes.byref(t0c), omega2,
                                    ctypes.cast(cb, ctypes.c_void_p) if cb is not None else None,
                                    1, ctypes.byref(err), ctypes.byref(te), cum)
    assert err.value == 0, lib.wendy_cuda_last_error()
    return t0c.value, te.value


@pytest.mark.parametrize('with_ext', [False, True])
def test_compat_export_same_signature_as_reference(with_ext):
    """Drive OUR library exactly as reference wendy/wendy.py:425-433 drives wendy_c."""
    from wendy_b200 import _lib
    lib = _lib.load()
    x0, v0, m = wo.sech2_ic(777, seed=12, mass_jitter=0.1)
    ext = (lambda x, t: -0.5 * numpy.tanh(x) + 0.01 * t) if with_ext else None
    ref = wo.COracle(x0, v0, m, 0.05, 4, t0=0.3, omega=1.1, ext_force=ext)
    x, v = x0.copy(), v0.copy()
    xi = numpy.zeros(len(x), dtype=_lib.XI_DTYPE)
    xi['idx'] = numpy.arange(len(x), dtype='i4')
    xi['val'] = x
    cb = wo.wrap_ext_force(ext)
    t0 = 0.3
    for _ in range(3):
        xr, vr = ref.step()
        t0, te = _compat_call(lib, xi, x, v, m, numpy.sum(m), 0.05 / 4, 4, t0, 1.1 ** 2, cb)
        assert relerr(x, xr) < 1e-12 and relerr(v, vr) < 1e-12
        assert numpy.array_equal(xi['idx'], ref.sidx)            # sort order of the last force
        assert numpy.array_equal(xi['val'], x[xi['idx']])        # wendy/wendy.c:413-415 invariant
        assert te > 0.
    assert t0 == (ref.t0.value if with_ext else 0.3)


# ---- adversarial inputs for the in-bucket interpolation sort and the splitter logic -----------------
def _nasty(kind, n, rs):
    if kind == 'duplicates':      # many exactly coincident particles (identical x AND v: they stay tied)
        x = numpy.repeat(rs.normal(size=n // 20), 20)
        v = numpy.repeat(rs.normal(size=n // 20), 20)
    elif kind == 'two_clumps':    # two tight clumps far apart + a thin background
        x = numpy.concatenate((rs.normal(-50., 1e-6, n // 2), rs.normal(80., 1e-3, n // 2 - 100), rs.uniform(-100, 100, 100)))
        v = rs.normal(size=n) * 0.01
    elif kind == 'heavy_tails':   # Cauchy positions: huge dynamic range inside edge buckets
        x = rs.standard_cauchy(size=n) * 1e3
        v = rs.standard_cauchy(size=n)
    elif kind == 'tiny_scale':    # everything within a few thousand ulps of 1.0
        x = 1.0 + rs.randint(0, 4000, size=n) * 2.0 ** -52
        v = rs.normal(size=n) * 1e-14
    elif kind == 'signed_zero':   # mixture of +0.0, -0.0 and denormals
        x = rs.choice([0.0, -0.0, 5e-324, -5e-324, 1e-310], size=n)
        v = rs.normal(size=n) * 1e-300
    return x, v


@pytest.mark.parametrize('cap', CAPS)
@pytest.mark.parametrize('kind', ['duplicates', 'two_clumps', 'heavy_tails', 'tiny_scale', 'signed_zero'])
def test_adversarial_distributions_bit_exact(kind, cap):
    import wendy_b200
    rs = numpy.random.RandomState(99)
    n = 4000
    x, v = _nasty(kind, n, rs)
    m = rs.uniform(0.5, 1.5, size=n) / n
    gen = wendy_b200.nbody(x, v, m, 0.03, approx=True, nleap=3, omega=0.3, _cap=cap)
    xo, vo = x, v
    for _ in range(2):
        xg, vg = next(gen)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.01, 3, 0.09, exact_scan=True)
        assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo), kind
    gen.close()


@pytest.mark.parametrize('by_id', ['', '1'])
@pytest.mark.parametrize('kind', ['duplicates', 'signed_zero', 'tiny_scale'])
def test_radix_path_orders_exact_ties_by_particle_index(kind, by_id, monkeypatch):
    """sort='gpu-radix' from the second sub-step on sorts keys taken in STORAGE order; the stable sort would leave
    exactly coincident keys in slot order, so the sorted values of every tie run are re-ordered by particle index
    (fix_ties_kernel: runs of 20 by insertion, runs of hundreds by heap sort).  Unequal masses make any wrong tie
    order visible in x, v.  WENDY_B200_RADIX_BYID=1: the earlier scheme (keys generated in id order), same answer."""
    import wendy_b200
    if by_id:
        monkeypatch.setenv('WENDY_B200_RADIX_BYID', by_id)
    rs = numpy.random.RandomState(98)
    n = 4000
    x, v = _nasty(kind, n, rs)
    m = rs.uniform(0.5, 1.5, size=n) / n
    gen = wendy_b200.nbody(x, v, m, 0.03, approx=True, nleap=3, omega=0.3, sort='gpu-radix')
    xo, vo = x, v
    for _ in range(3):
        xg, vg = next(gen)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.01, 3, 0.09, exact_scan=True)
        assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo), kind
    gen.close()


@pytest.mark.parametrize('jitter', [0., 0.2])
def test_launch_sequence_numbers_are_renewed_before_the_lookback_epoch_wraps(jitter, monkeypatch):
    """The look-back words keep 30 bits of the launch sequence number (round-1 advisor finding: a handle that had
    issued 2^30 launches would hang).  The handle starts again from 1 with clean look-back words long before that;
    here after every 40 launches, on the bucket path (count-prefix look-back; unequal masses: mass-prefix look-back
    too) and with an external force."""
    import torch
    import wendy_b200
    monkeypatch.setenv('WENDY_B200_EPOCH_RENEW_AT', '40')
    x, v, m = wo.sech2_ic(30000, seed=21, mass_jitter=jitter)
    st = wendy_b200.ApproxState(x, v, m, omega2=0.09)
    xo, vo = x, v
    for _ in range(12):
        st.step(0.01, 3)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.01, 3, 0.09, exact_scan=bool(jitter))
    F = lambda xx, t: -0.3 * xx  # noqa: E731
    t0 = 0.
    for _ in range(6):
        t0 = st.step_ext(0.01, 3, F, t0)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.01, 3, 0.09, F, 0., exact_scan=bool(jitter))
    xg, vg = st.read()
    st.close()
    assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)


@pytest.mark.parametrize('ext', [False, True])
def test_device_output_yields_the_same_numbers_without_leaving_the_gpu(ext):
    """nbody(..., output='device') yields CUDA tensors (de-sorted on the device): same values as the host arrays."""
    import torch
    import wendy_b200
    x, v, m = wo.sech2_ic(50000, seed=31)
    F = (lambda xx, t: -0.3 * xx + 0.01 * t) if ext else None  # noqa: E731
    gh = wendy_b200.nbody(x, v, m, 0.03, approx=True, nleap=3, omega=0.7, ext_force=F)
    gd = wendy_b200.nbody(x, v, m, 0.03, approx=True, nleap=3, omega=0.7, ext_force=F, output='device')
    for _ in range(3):
        xh, vh = next(gh)
        xd, vd = next(gd)
        assert xd.is_cuda and xd.dtype == torch.float64
        assert numpy.array_equal(xd.cpu().numpy(), xh) and numpy.array_equal(vd.cpu().numpy(), vh)
    gh.close(); gd.close()


@pytest.mark.parametrize('kind', ['steady', 'overflowing'])
def test_desort_ahead_of_the_readout_yields_the_same_outputs(kind, monkeypatch):
    """WENDY_B200_STAGE_AHEAD=1: the generator queues the de-sort of output k+1 behind the sub-steps of call k+1 (into
    a second staging set) while output k is still being copied.  Every yielded output must equal the oracle's --
    also when a sub-step of the call has to be re-run after a bucket overflow (the staged copy is then dropped)."""
    import wendy_b200
    monkeypatch.setenv('WENDY_B200_STAGE_AHEAD', '1')
    if kind == 'steady':
        x, v, m = wo.sech2_ic(30000, seed=33)
        kw, om, dtl = {}, 0.7, 0.01
    else:
        x, v, m = wo.slab_ic(20000, seed=5)
        kw, om, dtl = {'_cap': 256, '_fill': 250}, None, 0.05
    gen = wendy_b200.nbody(x, v, m, dtl * 5, approx=True, nleap=5, omega=om, **kw)
    xo, vo = x, v
    for _ in range(6):
        xg, vg = next(gen)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), dtl, 5, -1. if om is None else om ** 2.)
        assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo), kind
    gen.close()


def test_ext_force_on_an_ensemble_matches_separate_runs():
    """Config-5 shape in miniature: several realisations, torch-vectorised external force."""
    import torch
    import wendy_b200
    S, L = 4, 3000
    F = lambda x, t: -0.7 * torch.tanh(0.5 * x) + 0.01 * t  # noqa: E731
    Fn = lambda x, t: -0.7 * numpy.tanh(0.5 * x) + 0.01 * t  # noqa: E731
    ics = [wo.sech2_ic(L, seed=40 + s) for s in range(S)]
    X = numpy.concatenate([i[0] for i in ics]); V = numpy.concatenate([i[1] for i in ics])
    M = numpy.concatenate([0.3 * i[2] for i in ics])
    gen = wendy_b200.nbody(X, V, M, 0.05, approx=True, nleap=5, ext_force=F, t0=0.2, n_segments=S)
    for _ in range(2):
        Xg, Vg = next(gen)
    gen.close()
    for s in range(S):
        xo, vo, t0 = ics[s][0], ics[s][1], 0.2
        for _ in range(2):
            xo, vo, t0, _ = wo.numpy_onestep(xo, vo, 0.3 * ics[s][2], numpy.sum(0.3 * ics[s][2]), 0.01, 5, -1., Fn, t0)
        # torch.tanh and numpy.tanh may differ in the last bit: tolerance instead of bit equality
        assert relerr(Xg[s * L:(s + 1) * L], xo) < 1e-12 and relerr(Vg[s * L:(s + 1) * L], vo) < 1e-12


def test_device_constructor_and_ic_generators_match_host_path():
    """from_device (no host staging) must give exactly what the host constructor gives for the same data."""
    import torch
    import wendy_b200
    from wendy_b200 import ic
    for gen_ic in (ic.sech2_disk, ic.cold_slab, ic.exponential_disk):
        x, v, m0 = gen_ic(50000, seed=5)
        a = wendy_b200.ApproxState.from_device(x, v, m0, omega2=0.25)
        xh, vh = x.cpu().numpy(), v.cpu().numpy()
        b = wendy_b200.ApproxState(xh, vh, numpy.full(len(xh), m0), omega2=0.25)
        b_tot = numpy.sum(numpy.full(len(xh), m0))
        for st in (a, b):
            st.step(0.01, 5)
        xa, va = a.read(); xb, vb = b.read()
        a.close(); b.close()
        if b_tot == m0 * len(xh):  # identical total mass -> identical results
            assert numpy.array_equal(xa, xb) and numpy.array_equal(va, vb)
        else:
            assert relerr(xa, xb) < 1e-13 and relerr(va, vb) < 1e-13
    # general masses from a device tensor
    x, v, m0 = ic.sech2_disk(20000, seed=6)
    m = torch.full_like(x, m0) * (1. + 0.1 * torch.sin(torch.arange(len(x), device=x.device, dtype=torch.float64)))
    a = wendy_b200.ApproxState.from_device(x, v, m)
    b = wendy_b200.ApproxState(x.cpu().numpy(), v.cpu().numpy(), m.cpu().numpy())
    for st in (a, b):
        st.step(0.01, 3)
    xa, va = a.read(); xb, vb = b.read()
    a.close(); b.close()
    assert relerr(xa, xb) < 1e-13 and relerr(va, vb) < 1e-13


# ---- robustness ------------------------------------------------------------------------------------
def test_overflow_recovery_by_rebalancing():
    """A violently collapsing cold slab changes the density by orders of magnitude: buckets
    overflow, the step is re-run after re-balancing, and the result still matches the oracle."""
    import wendy_b200
    x, v, m = wo.slab_ic(20000, seed=5)
    st = wendy_b200.ApproxState(x, v, m, cap=256, fill=250)
    xo, vo = x, v
    for _ in range(6):
        st.step(0.05, 5)
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.05, 5)  # equal masses: serial scan
    xg, vg = st.read()
    s = st.stats()
    st.close()
    assert s['failed_substeps'] > 0 and s['rebuilds'] > 1, s
    assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)


@pytest.mark.parametrize('nleap,calls,async_ext', [(1, 8, '1'), (4, 2, '1'), (4, 2, '0')])
def test_ext_force_with_overflow_on_the_first_substep_applies_the_half_drift_once(nleap, calls, async_ext, monkeypatch):
    """A bucket overflow on the first sub-step of an output step makes the library restore that sub-step's input
    after the leading half drift was materialised; the retry must not drift again (round-1 advisor finding).
    nleap = 4: the overflow hits a later sub-step of a call that was enqueued asynchronously (wendy_cuda_ext_begin /
    _substep_async / _ext_end): the launches behind it are void and the call is finished synchronously."""
    import wendy_b200
    monkeypatch.setenv('WENDY_B200_EXT_ASYNC', async_ext)
    x, v, m = wo.slab_ic(20000, seed=5)
    F = lambda xx, t: -0.3 * xx + 0.05 * t  # noqa: E731  (linear: bit-identical in torch and numpy)
    st = wendy_b200.ApproxState(x, v, m, cap=256, fill=250)
    xo, vo, t0, tg = x, v, 0.1, 0.1
    for _ in range(calls):
        tg = st.step_ext(0.05, nleap, F, tg)
        xo, vo, t0, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.05, nleap, -1., F, t0)
    xg, vg = st.read()
    s = st.stats()
    st.close()
    assert s['failed_substeps'] > 0, s
    assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)


def test_optimistic_fill_backs_off_after_an_overflow():
    """Large equal-mass systems start on 2048-slot buckets filled to 13/16; the first overflow (here: a
    collapsing cold slab) moves the handle back to 3/4 for good, and overflows that keep coming buy sparser
    buckets still (half the fill, as far as the storage goes).  Results equal the radix path bit for bit."""
    import wendy_b200
    n = 1 << 20
    x, v, m = wo.slab_ic(n, seed=8)
    a = wendy_b200.ApproxState(x, v, m)
    b = wendy_b200.ApproxState(x, v, m, sort='gpu-radix')
    a.step(0.001, 2)
    b.step(0.001, 2)
    s0 = a.stats()
    assert s0['cap'] == 2048 and s0['buckets'] == -(-n // 1664), s0
    for _ in range(6):
        a.step(0.05, 5)
        b.step(0.05, 5)
    s1 = a.stats()
    xa, va = a.read(); xb, vb = b.read()
    a.close(); b.close()
    assert s1['failed_substeps'] > 0 and s1['buckets'] >= -(-n // 1536), s1
    assert numpy.array_equal(xa, xb) and numpy.array_equal(va, vb)


def test_adaptive_layout_switches_to_coarse_buckets_and_stays_exact():
    """With a large N*dt most particles leave the 32-bucket window of the warp kernel; the library
    then rebuilds the layout with 2048-slot buckets (CTA kernel).  Results must not change."""
    import wendy_b200
    x, v, m = wo.sech2_ic(400000, seed=13)
    st = wendy_b200.ApproxState(x, v, m, omega2=1.21)
    xo, vo = x, v
    caps = []
    for _ in range(3):
        st.step(0.1, 3)
        caps.append(st.stats()['cap'])
        xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.1, 3, 1.21)  # equal masses: serial scan
    xg, vg = st.read()
    st.close()
    assert caps[0] in (256, 2048) and caps[-1] == 2048, caps
    assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)


@pytest.mark.parametrize('grid', ['1', '3'])
@pytest.mark.parametrize('ext', [False, True])
def test_persistent_kernel_many_buckets_per_cta(monkeypatch, grid, ext):
    """The persistent CTA kernel walks several buckets per CTA with TMA prefetch, double-buffered counters
    and splitter windows; a tiny grid makes every CTA iterate ~10-30 times on a small system.  Bit-exact
    against the serial-scan oracle (equal masses; plain instance and the instance with an external-force array)."""
    import wendy_b200
    monkeypatch.setenv('WENDY_B200_PERSIST_GRID', grid)
    x, v, m = wo.sech2_ic(50000, seed=21)
    kw = dict(omega=1.1)
    if ext:
        kw = dict(ext_force=lambda xx, t: -1.21 * xx + 0.1 * t, t0=0.5)
    g = wendy_b200.nbody(x, v, m, 0.02, approx=True, nleap=4, _cap=2048, **kw)
    xo, vo = x, v
    t0 = 0.5
    for _ in range(3):
        xg, vg = next(g)
        if ext:
            xo, vo, t0, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.005, 4, -1.,
                                             ext_force=lambda xx, t: -1.21 * xx + 0.1 * t, t0=t0)
        else:
            xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), 0.005, 4, 1.1 ** 2.)
        assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)
    g.close()


def test_rejects_non_finite_input():
    import wendy_b200
    with pytest.raises(RuntimeError):
        wendy_b200.ApproxState(numpy.array([0., numpy.nan]), numpy.zeros(2), numpy.ones(2))

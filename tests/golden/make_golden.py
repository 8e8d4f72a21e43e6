"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run here (the dev container), never on the GPU box:

    make -C oracle ref && python tests/golden/make_golden.py

Two routes into the reference are used, and cross-checked against each other:
  (1) ``import wendy`` from /root/reference (its Python generator + numba/ctypes
      ext_force adapters + wendy.energy), with oracle/_ref/wendy_c.so exposed under the
      name the reference's loader looks for (wendy/wendy.py:21-29);
  (2) oracle.wendy_oracle.Reference: the reference's C entry point through ctypes.
Inputs are seeded or literal, so the vectors are reproducible.
"""
import os
import shutil
import sys
import sysconfig
import tempfile

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('OMP_NUM_THREADS', '1')
from oracle import wendy_oracle as wo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tmp = tempfile.mkdtemp(prefix='wendy_ref_')
    shutil.copy(os.path.join(ROOT, 'oracle', '_ref', 'wendy_c.so'),
                os.path.join(tmp, 'wendy_c%s' % sysconfig.get_config_var('EXT_SUFFIX')))
    sys.path.insert(0, tmp)
    sys.path.insert(0, '/root/reference')
    import wendy
    return wendy


def run_py(wendy, x, v, m, dt, nleap, nout, **kw):
    g = wendy.nbody(x, v, m, dt, approx=True, nleap=nleap, **kw)
    xs, vs = [], []
    for _ in range(nout):
        tx, tv = next(g)
        xs.append(tx.copy())
        vs.append(tv.copy())
    return numpy.array(xs), numpy.array(vs)


def run_c(x, v, m, dt, nleap, nout, **kw):
    r = wo.Reference(x, v, m, dt, nleap, **kw)
    xs, vs = [], []
    for _ in range(nout):
        tx, tv = r.step()
        xs.append(tx.copy())
        vs.append(tv.copy())
    return numpy.array(xs), numpy.array(vs), r


def main():
    wendy = import_reference()
    cases = {}

    # --- KAT-A/B/C: three bodies -------------------------------------
    x = numpy.array([-1.1, 0.1, 1.3])
    v = numpy.array([3., 2., -5.])
    m = numpy.array([1., 2., 3.])
    for name, kw in [('kat_a', {}), ('kat_b', {'omega': 1.1}),
                     ('kat_c', {'ext_force': (lambda x, t: -1.21 * x + 0.1 * t), 't0': 0.5})]:
        xs, vs = run_py(wendy, x, v, m, 0.05, 10, 3, **kw)
        xc, vc, _ = run_c(x, v, m, 0.05, 10, 3, **kw)
        assert numpy.array_equal(xs, xc) and numpy.array_equal(vs, vc), name
        om = kw.get('omega', None)
        E = numpy.array([wendy.energy(x, v, m, omega=om)] +
                        [wendy.energy(xs[i], vs[i], m, omega=om) for i in range(3)])
        cases[name] = dict(x0=x, v0=v, m=m, dt=0.05, nleap=10, xs=xs, vs=vs, E=E,
                           omega=numpy.nan if om is None else om, t0=kw.get('t0', 0.))

    # --- sech^2 disk with +-10 % mass jitter (reference tests/test_approx.py:34-51) --
    x, v, m = wo.sech2_ic(1000, seed=2, mass_jitter=0.1)
    for name, nleap, nout, kw in [
            ('sech2_1000_nleap1', 1, 10, {}),
            ('sech2_1000_nleap7_omega', 7, 4, {'omega': 1.1}),
            ('sech2_1000_twopiG', 2, 4, {'twopiG': 2.5}),
            ('sech2_1000_ext', 5, 4, {'ext_force': (lambda x, t: -0.7 * numpy.tanh(0.5 * x) + 0.05 * t),
                                      't0': 0.25})]:
        for srt in ('merge', 'quick', 'parallel'):
            xc, vc, r = run_c(x, v, m, 0.05, nleap, nout, sort=srt, **kw)
            if srt == 'merge':
                xs, vs, ids = xc, vc, r.xi['idx'].copy()
            else:  # all reference sorts agree for distinct keys
                assert numpy.array_equal(xs, xc) and numpy.array_equal(vs, vc), (name, srt)
        if 'ext_force' not in kw:
            xp, vp = run_py(wendy, x, v, m, 0.05, nleap, nout, sort='merge', **kw)
            assert numpy.array_equal(xs, xp) and numpy.array_equal(vs, vp), name
        tg = kw.get('twopiG', 1.)
        E = numpy.array([wendy.energy(x, v, m, twopiG=tg, omega=kw.get('omega'))] +
                        [wendy.energy(xs[i], vs[i], m, twopiG=tg, omega=kw.get('omega'))
                         for i in range(nout)])
        cases[name] = dict(x0=x, v0=v, m=m, dt=0.05, nleap=nleap, xs=xs, vs=vs, E=E,
                           last_sorted_ids=ids, omega=kw.get('omega', numpy.nan),
                           twopiG=tg, t0=kw.get('t0', 0.))

    # --- KAT-D / config 1: N=1e4 sech^2, dt=0.05, nleap=1, 100 outputs ----------------
    numpy.random.seed(2)
    N = 10000
    x = numpy.arctanh(2. * numpy.random.uniform(size=N) - 1) * 2.
    v = numpy.random.normal(size=N)
    v -= numpy.mean(v)
    m = numpy.ones(N) / N
    xc, vc, r = run_c(x, v, m, 0.05, 1, 100, sort='merge')
    keep = [0, 9, 99]
    E = numpy.array([wendy.energy(x, v, m)] + [wendy.energy(xc[i], vc[i], m) for i in keep])
    cases['config1_sech2_1e4'] = dict(x0=x, v0=v, m=m, dt=0.05, nleap=1, keep=numpy.array(keep),
                                      xs=xc[keep], vs=vc[keep], E=E,
                                      last_sorted_ids=r.xi['idx'].copy())

    # --- config 2 in miniature: cold slab collapse, N=4096 -----------------------------
    x, v, m = wo.slab_ic(4096, seed=3)
    xc, vc, r = run_c(x, v, m, 0.05, 10, 40, sort='merge')
    keep = [0, 9, 39]
    cases['slab_4096'] = dict(x0=x, v0=v, m=m, dt=0.05, nleap=10, keep=numpy.array(keep),
                              xs=xc[keep], vs=vc[keep], last_sorted_ids=r.xi['idx'].copy())

    # --- coincident particles at the first force evaluation (tests/test_approx.py:234-248)
    # 'merge' is stable w.r.t. the initial (index) order, so this pins "ties by index".
    x = numpy.array([-1.1, 0.1, 0.1, 1.3, 0.1, -1.1, 2.])
    v = numpy.array([3., 2., 2., -5., 2., 3., 0.5])  # equal v: the tie survives the half drift
    m = numpy.array([1., 2., 3., 1., 0.5, 2., 1.5])
    # ONE output only: these particles re-coincide later by symmetry, and there the
    # reference's stable sorts keep the *previous array order* while this framework's
    # convention is (key, particle index) -- see SURVEY.md "three things" #2.
    xc, vc, r = run_c(x, v, m, 0.05, 3, 1, sort='merge')
    cases['ties'] = dict(x0=x, v0=v, m=m, dt=0.05, nleap=3, xs=xc, vs=vc,
                         last_sorted_ids=r.xi['idx'].copy())

    # --- tracer (zero-mass) particles (tests/test_approx.py:165-185) -------------------
    x, v, m = wo.sech2_ic(300, seed=5, mass_jitter=0.1)
    m[::3] = 0.
    xc, vc, r = run_c(x, v, m, 0.05, 4, 5, sort='merge')
    cases['tracers'] = dict(x0=x, v0=v, m=m, dt=0.05, nleap=4, xs=xc, vs=vc)

    # --- physics cross-check against the reference's EXACT event-driven solver ------------------
    # (reference tests/test_approx.py:213-232: approx with nleap=2000 tracks the exact solution to 1e-5)
    rs = numpy.random.RandomState(2)
    N = 101
    x = numpy.arctanh(2. * rs.uniform(size=N) - 1) * 2.
    v = rs.normal(size=N)
    v -= numpy.mean(v)
    m = numpy.ones(N) / N * (1. + 0.1 * (2. * rs.uniform(size=N) - 1))
    for name, om in (('exact_solver_101', None), ('exact_solver_101_harm', 1.1)):
        g = wendy.nbody(x, v, m, 0.05, omega=om)  # approx=False: wendy/wendy.c:53-315 (out of scope here)
        xs, vs = [], []
        for _ in range(20):
            tx, tv = next(g)
            xs.append(tx.copy())
            vs.append(tv.copy())
        cases[name] = dict(x0=x, v0=v, m=m, dt=0.05, xs=numpy.array(xs), vs=numpy.array(vs),
                           omega=numpy.nan if om is None else om)

    # --- diagnostics: potential(y) and per-particle energies (wendy/wendy.py:494-517, 466-470) -----------
    rs = numpy.random.RandomState(7)
    N = 1000
    x = numpy.arctanh(2. * rs.uniform(size=N) - 1) * 2.
    x[100:110] = x[100]  # coincident particles
    v = rs.normal(size=N)
    m = numpy.ones(N) / N * (1. + 0.5 * (2. * rs.uniform(size=N) - 1))
    y = numpy.concatenate((numpy.linspace(-12., 12., 201), x[::50], [x.min(), x.max(), -50., 50.]))
    cases['potential_1000'] = dict(
        x=x, v=v, m=m, y=y, twopiG=1.7, omega=1.1,
        pot=wendy.potential(y, x, v, m, twopiG=1.7), pot_harm=wendy.potential(y, x, v, m, twopiG=1.7, omega=1.1),
        eind=wendy.energy(x, v, m, twopiG=1.7, individual=True),
        eind_harm=wendy.energy(x, v, m, twopiG=1.7, individual=True, omega=1.1))

    only = sys.argv[1:]
    for name, d in cases.items():
        if only and name not in only:
            continue
        numpy.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
        print('wrote', name, {k: numpy.shape(a) for k, a in d.items()})
    import hashlib
    d = cases['config1_sech2_1e4']
    print('KAT-D sha256(x)[:16] =', hashlib.sha256(d['xs'][-1].tobytes()).hexdigest()[:16],
          ' sha256(v)[:16] =', hashlib.sha256(d['vs'][-1].tobytes()).hexdigest()[:16])


if __name__ == '__main__':
    main()

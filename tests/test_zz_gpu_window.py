"""GPU: the destination window of the persistent CTA kernel (wendy_b200/csrc/tile.cu) --
its width, pinned or adapted by the library, must never change a result.  (Last file of the GPU suite by name.)"""
import numpy
import pytest

from oracle import wendy_oracle as wo

pytestmark = pytest.mark.gpu

_WINDOW_ORACLE = {}


@pytest.mark.parametrize('dw', [None, '8', '64', '1024'])
def test_destination_window_width_does_not_change_results(monkeypatch, dw):
    """The persistent CTA kernel keeps the splitters of a window of buckets around the home bucket in shared
    memory; new keys outside it take a galloping search in the global table, keys far from home inside a wide
    window a secant refinement of the interpolated guess.  The library widens the window (256 -> 1024 buckets)
    when particles keep leaving it; WENDY_B200_DW pins it.  1300 buckets and sub-steps that carry particles
    across tens to hundreds of them: every width must give the oracle's result bit for bit (equal masses)."""
    import wendy_b200
    if dw is None:
        monkeypatch.delenv('WENDY_B200_DW', raising=False)
    else:
        monkeypatch.setenv('WENDY_B200_DW', dw)
    n, dt, nleap, calls = 2000000, 0.3, 3, 2  # per sub-step: median shift 15 buckets, 1 % beyond 80, tail beyond 128
    x, v, m = wo.sech2_ic(n, seed=31)
    if 'xv' not in _WINDOW_ORACLE:
        xo, vo = x, v
        for _ in range(calls):
            xo, vo, _, _ = wo.numpy_onestep(xo, vo, m, numpy.sum(m), dt / nleap, nleap, 1.1 ** 2.)
        _WINDOW_ORACLE['xv'] = (xo, vo)
    xo, vo = _WINDOW_ORACLE['xv']
    g = wendy_b200.nbody(x, v, m, dt, approx=True, nleap=nleap, omega=1.1, _cap=2048)
    for _ in range(calls):
        xg, vg = next(g)
    g.close()
    assert numpy.array_equal(xg, xo) and numpy.array_equal(vg, vo)

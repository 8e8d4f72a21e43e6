"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy
sys.path.insert(0, '.')
import wendy_b200
from oracle import wendy_oracle as wo
for cap, general in ((0, False), (0, True), (2048, False), (2048, True)):
    x, v, m = wo.sech2_ic(6000, seed=3, mass_jitter=0.1 if general else 0.)
    for sort in ('gpu', 'gpu-radix'):
        g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=3, omega=1.1, sort=sort, _cap=cap)
        for _ in range(2):
            xg, vg = next(g)
        g.close()
    print('cap', cap, 'general', general, 'E', wendy_b200.energy(x, v, m, omega=1.1), flush=True)
# persistent CTA kernel with several buckets per CTA (grid shrunk to 2 CTAs)
import os
os.environ['WENDY_B200_PERSIST_GRID'] = '2'
x, v, m = wo.sech2_ic(20000, seed=4)
for kw in (dict(omega=1.1), dict(ext_force=lambda xx, t: -1.21 * xx)):
    g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=3, _cap=2048, **kw)
    for _ in range(2):
        xg, vg = next(g)
    g.close()
print('persistent, 2 CTAs: ok', flush=True)
print('done')

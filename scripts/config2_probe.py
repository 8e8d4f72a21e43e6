"""Config 2 (cold slab N=1e6, dt 0.005, 1000 single-sub-step calls) under several layout settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import wendy_b200
from bench import slab_ic
nn, nsteps, dtl = 1000000, 1000, 0.005
x, v, m = slab_ic(nn)
for label, kw, nleap in (('default', {}, 1), ('cap2048 fill1024', {'cap': 2048, 'fill': 1024}, 1), ('cap2048 fill768', {'cap': 2048, 'fill': 768}, 1),
                         ('cap256', {'cap': 256}, 1), ('cap256 fill96', {'cap': 256, 'fill': 96}, 1), ('radix', {'sort': 'gpu-radix'}, 1),
                         ('default nleap=10', {}, 10)):
    st = wendy_b200.ApproxState(x, v, m, **kw)
    st.step(dtl, nleap)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(nsteps // nleap - 1):
        st.step(dtl, nleap)
    torch.cuda.synchronize()
    el = time.perf_counter() - t
    s = st.stats()
    print('%-18s %.3e particle-steps/s  %.1f us/sub-step  rebuilds %d failed %d radix %d launches %d cap %d buckets %d' % (
        label, nn * (nsteps - nleap) / el, 1e6 * el / (nsteps - nleap), s['rebuilds'], s['failed_substeps'], s['radix_fallbacks'],
        s['kernel_launches'], s['cap'], s['buckets']), flush=True)
    st.close()

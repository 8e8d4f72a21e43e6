"""Which combination makes the coarse (2048-slot) path fail every sub-step: segments x ext_force x geometry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import wendy_b200
from oracle import wendy_oracle as wo

F = lambda xx, t: -0.7 * torch.tanh(0.5 * xx)  # noqa: E731
for nseg, L, cap, fill, ext in ((4, 300000, 0, 0, False), (4, 300000, 2048, 0, True), (4, 300000, 2048, 1664, True), (4, 300000, 2048, 1664, False),
                                (4, 300000, 0, 0, True), (3, 400000, 0, 0, True), (2, 600000, 0, 0, True)):
    xs, vs = numpy.empty(nseg * L), numpy.empty(nseg * L)
    for j in range(nseg):
        x, v, m = wo.sech2_ic(L, seed=3 + j)
        xs[j * L:(j + 1) * L], vs[j * L:(j + 1) * L] = x, v
    ms = numpy.full(nseg * L, 0.3 / L)
    st = wendy_b200.ApproxState(xs, vs, ms, n_segments=nseg, cap=cap, fill=fill)
    t0 = 0.
    for c in range(3):
        if ext:
            t0 = st.step_ext(0.005, 4, F, t0)
        else:
            st.step(0.005, 4)
    s = st.stats()
    print('nseg %d L %d cap %d fill %d ext %s ->' % (nseg, L, cap, fill, ext), {k: s[k] for k in ('substeps', 'rebuilds', 'failed_substeps', 'radix_fallbacks', 'cap', 'buckets', 'max_bucket_count')}, flush=True)
    st.close()

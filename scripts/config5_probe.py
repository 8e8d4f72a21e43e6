"""BASELINE config 5 shape on one GPU, call by call: where the time of the ext_force path goes (layout statistics,
time per call with and without the read-out).  python scripts/config5_probe.py [realisations] [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import torch

import wendy_b200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = 100000
alpha, sigma, zh = 0.3, 1., 1.
xs, vs = numpy.empty(S * L), numpy.empty(S * L)
for j in range(S):
    rs = numpy.random.RandomState(2 + j)
    xs[j * L:(j + 1) * L] = numpy.arctanh(2. * rs.uniform(size=L) - 1.) * 2. * zh
    vv = rs.normal(size=L) * sigma
    vs[j * L:(j + 1) * L] = vv - numpy.mean(vv) + sigma
ms = numpy.full(S * L, alpha / L)
F = lambda xx, t: -(1. - alpha) * sigma ** 2. * torch.tanh(0.5 * xx / zh) / zh  # noqa: E731
for kick in (sigma, 0.):
    if kick == 0.:
        vs -= sigma
    st = wendy_b200.ApproxState(xs, vs, ms, n_segments=S)
    t0 = 0.
    prev = st.stats()
    for c in range(calls):
        torch.cuda.synchronize()
        t = time.perf_counter()
        t0 = st.step_ext(0.005, 10, F, t0)
        torch.cuda.synchronize()
        el = time.perf_counter() - t
        s = st.stats()
        print('bulk kick %.0f call %d: %.1f ms for 10 sub-steps (%.2e particle-steps/s)  rebuilds +%d failed +%d radix +%d '
              'launches +%d cap %d buckets %d left_window +%d' % (
                  kick, c, 1e3 * el, S * L * 10 / el, s['rebuilds'] - prev['rebuilds'],
                  s['failed_substeps'] - prev['failed_substeps'], s['radix_fallbacks'] - prev['radix_fallbacks'],
                  s['kernel_launches'] - prev['kernel_launches'], s['cap'], s['buckets'],
                  s['left_window'] - prev['left_window']), flush=True)
        prev = s
    st.close()

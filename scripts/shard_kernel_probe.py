"""One GPU: time the sub-step through the shard entry points (nranks=1, no migrants) against the plain handle."""
import os
import sys
import time

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wendy_b200
from wendy_b200 import multi
from bench import sech2_ic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
dt = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
x, v, m = sech2_ic(n, 2)
ids = numpy.arange(n, dtype=numpy.int32)
m0 = 1. / n
bounds = numpy.array([-numpy.inf, numpy.inf])
eng = multi.CudaShardEngine(x, v, ids, m0, 1.0, 1.21, 1, 0, bounds, int(1.3 * n), 1 << 20)
for rep in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for k in range(10):
        eng.substep(dt / 2. if k == 0 else 0., dt, dt / 2. if k == 9 else dt, dt / 2. if k == 9 else 0., 0)
    torch.cuda.synchronize()
    print('shard handle: %.3f ms per sub-step' % ((time.perf_counter() - t) * 100.))
eng.close()
st = wendy_b200.ApproxState(x, v, numpy.full(n, m0), omega2=1.21)
for rep in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    st.step(dt, 10)
    torch.cuda.synchronize()
    print('plain handle: %.3f ms per sub-step' % ((time.perf_counter() - t) * 100.), st.stats()['cap'])
st.close()

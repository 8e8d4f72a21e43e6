"""Where the end-to-end time of the generator API goes at N=1e8 (host buffers): phase timings."""
import os
import sys
import time

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wendy_b200
from wendy_b200 import _lib
from bench import sech2_ic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
x0, v0, m = sech2_ic(n, 2)
torch.cuda.init()
lib = _lib.load()


def T(label, t):
    torch.cuda.synchronize()
    print('%-44s %8.1f ms' % (label, 1e3 * (time.perf_counter() - t)), flush=True)


t = time.perf_counter(); x = numpy.array(x0); v = numpy.array(v0); T('numpy copies of x, v', t)
t = time.perf_counter(); lib.wendy_cuda_pin(x.ctypes.data, x.nbytes); lib.wendy_cuda_pin(v.ctypes.data, v.nbytes); T('page-lock x, v (cudaHostRegister)', t)
t = time.perf_counter(); st = wendy_b200.ApproxState(x, v, m, omega2=1.21); T('ApproxState (validate, alloc, H2D)', t)
t = time.perf_counter(); st.step(1e-4, 10); T('first call (layout build + 10 sub-steps)', t)
t = time.perf_counter(); st.step(1e-4, 10); T('second call (10 sub-steps)', t)
t = time.perf_counter(); st.read(x, v); T('read (unsort + D2H into pinned x, v)', t)
st.close()
lib.wendy_cuda_unpin(x.ctypes.data); lib.wendy_cuda_unpin(v.ctypes.data)
t = time.perf_counter()
g = wendy_b200.nbody(x0, v0, m, 1e-3, approx=True, nleap=10, omega=1.1)
next(g); T('generator: construction + first output', t)
t = time.perf_counter()
for _ in range(5):
    next(g)
T('generator: 5 more outputs', t)
g.close()

"""A coherently drifting disk (bulk velocity u): fixed bucket edges see the density slide past them, advected
edges move with it.  Reports throughput and layout rebuilds; run with WENDY_B200_ADVECT=0/1."""
import os, sys, time
import numpy
sys.path.insert(0, '.')
import torch
import wendy_b200
from oracle import wendy_oracle as wo
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4000000
u = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
x, v, m = wo.sech2_ic(n, seed=3)
v = v + u
st = wendy_b200.ApproxState(x, v, m)
st.step(1e-3, 10)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(40):
    st.step(1e-3, 10)
torch.cuda.synchronize(); el = time.perf_counter() - t0
print('ADVECT=%s N=%d u=%g: %.3e particle-steps/s' % (os.environ.get('WENDY_B200_ADVECT', '1'), n, u, n * 400 / el), st.stats())
st.close()

"""BASELINE config 2 in full: cold-slab phase mixing / violent relaxation, N=1e6, omega=0, 1000 leapfrog
steps of dt_leap=0.005 on one B200; reports throughput, the path statistics and the energy drift."""
import sys, time
import numpy
sys.path.insert(0, '.')
import torch
import wendy_b200
from oracle import wendy_oracle as wo
x, v, m = wo.slab_ic(1000000, seed=3)
E0 = wendy_b200.energy(x, v, m, omega=0.)
st = wendy_b200.ApproxState(x, v, m, omega2=0.)
st.step(0.005, 10)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(99):
    st.step(0.005, 10)
torch.cuda.synchronize(); el = time.perf_counter() - t0
xg, vg = st.read()
print('config 2: N=1e6, 990 sub-steps in %.3f s -> %.3e particle-steps/s' % (el, 1e6 * 990 / el))
print('stats', st.stats())
print('dE/E after 1000 steps: %.3e' % ((wendy_b200.energy(xg, vg, m, omega=0.) - E0) / E0))
st.close()

"""Every kernel of the path once, for a per-kernel ncu pass (profiles/r02/kernels.md):
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv python scripts/kernel_tour.py
Sections are separated by NVTX-free markers: the summariser only groups by kernel name and launch order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import torch

import wendy_b200
from wendy_b200 import ic

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000

# 1. large equal-mass system, device ICs: layout build (make_keys, onesweep radix, pick_splitters, scatter), the
#    persistent step kernel (+ count prefix), energy (radix + gather tile kernel + reduce), read-out (unsort)
x, v, m0 = ic.sech2_disk(N, seed=2)
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21)
st.step(1e-3, 3)
print('energy terms', st.energy_terms(), flush=True)
xo = torch.empty(N, dtype=torch.float64).numpy()
vo = torch.empty(N, dtype=torch.float64).numpy()
st.read(xo, vo)
st.close()
# 2. the same on the warp kernel (256-slot buckets) and with the radix sort forced every sub-step
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21, cap=256)
st.step(1e-5, 2)
st.close()
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21, sort='gpu-radix')
st.step(1e-3, 1)
st.close()
# 3. device potential(y) and per-particle energies (potential.cu)
n_p = min(N, 20000000)
xs, vs = x[:n_p].contiguous(), v[:n_p].contiguous()
ms = torch.full((n_p,), 1. / n_p, dtype=torch.float64, device='cuda')
y = torch.linspace(-3., 3., 100001, dtype=torch.float64, device='cuda')
wendy_b200.potential(y, xs, vs, ms, omega=1.1)
wendy_b200.energy(xs, vs, ms, individual=True, omega=1.1)
del x, v
# 4. general (unequal) masses: mass prefix kernels + the general instances of the warp / CTA kernels
n_g = min(N, 20000000)
rs = numpy.random.RandomState(5)
xg = numpy.arctanh(2. * rs.uniform(size=n_g) - 1.) * 2.
vg = rs.normal(size=n_g)
mg = (1. + 0.1 * rs.uniform(size=n_g)) / n_g
for cap in (0, 2048):
    st = wendy_b200.ApproxState(xg, vg, mg, omega2=1.21, cap=cap)
    st.step(1e-4, 2)
    st.close()
# 5. an ensemble of small systems (small_kernel: one CTA per system, all sub-steps in one launch)
ne, npart = 2000, 1000
xe = numpy.arctanh(2. * rs.uniform(size=ne * npart) - 1.) * 2.
ve = rs.normal(size=ne * npart)
me = numpy.full(ne * npart, 1. / npart)
st = wendy_b200.ApproxState(xe, ve, me, omega2=1.21, n_segments=ne)
st.step(1e-3, 100)
st.close()
# 6. ensemble of mid-size systems with a torch ext_force (config 5 shape, reduced): force_positions / apply_drift
ne, npart = 64, 100000
xe = numpy.arctanh(2. * rs.uniform(size=ne * npart) - 1.) * 2.
ve = rs.normal(size=ne * npart)
me = numpy.full(ne * npart, 1. / npart)
g = wendy_b200.nbody(xe, ve, me, 0.01, approx=True, nleap=2, n_segments=ne, ext_force=lambda xx, t: -0.1 * torch.tanh(xx))
next(g)
g.close()
print('tour done', flush=True)

"""N = 1e9 particles on ONE B200 (device-generated sech^2 disk, equal masses): memory and throughput check."""
import sys, time
sys.path.insert(0, '.')
import torch
import wendy_b200
from wendy_b200 import ic
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000000
dt = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
x, v, m0 = ic.sech2_disk(n, seed=2)
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21)
del x, v
torch.cuda.empty_cache()
st.step(dt, 2)
torch.cuda.synchronize(); t0 = time.perf_counter()
st.step(dt, 10)
torch.cuda.synchronize(); el = time.perf_counter() - t0
free, total = torch.cuda.mem_get_info()
print('N=%d dt_leap=%g: %.3e particle-steps/s, %.1f ms per sub-step, device memory in use %.1f GB of %.1f'
      % (n, dt, n * 10 / el, el * 100, (total - free) / 1e9, total / 1e9), st.stats())
e = st.energy_terms()
print('energy terms (kinetic, harmonic, potential, momentum):', e)
st.close()

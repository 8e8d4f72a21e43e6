"""Time the radix sort alone (through wendy_b200.argsort's device path is host-bound; use the forced-radix step and the layout build)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wendy_b200
from wendy_b200 import ic
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
x, v, m0 = ic.sech2_disk(n, seed=2)
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21, sort='gpu-radix')
st.step(1e-3, 2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
st.step(1e-3, 6)
e1.record(); torch.cuda.synchronize()
print('%s forced radix: %.3f ms per sub-step' % (os.environ.get('WENDY_B200_LIB', 'default'), e0.elapsed_time(e1) / 6), flush=True)
st.close()

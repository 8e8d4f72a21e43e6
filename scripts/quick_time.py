"""Smallest possible timing of one library build: N=1e8 device ICs, dt_leap=1e-3, 2 warm-up + 3 timed calls."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wendy_b200
from wendy_b200 import ic
x, v, m0 = ic.sech2_disk(3 << 20, seed=7)
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21)
st.step(1e-3, 10); st.step(1e-3, 10)
xo, vo = st.read(); st.close()
print('fingerprint', hashlib.sha256(xo.tobytes() + vo.tobytes()).hexdigest()[:16], flush=True)
x, v, m0 = ic.sech2_disk(100000000, seed=2)
st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21)
for _ in range(2):
    st.step(1e-3, 10)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    st.step(1e-3, 10)
e1.record(); torch.cuda.synchronize()
print('value %.4e' % (1e8 * 30 / (e0.elapsed_time(e1) * 1e-3)), st.stats(), flush=True)

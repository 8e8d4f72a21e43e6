"""Timing of the device diagnostics at N=1e8 (potential on 1e5 points, per-particle energies)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import wendy_b200
from wendy_b200 import ic

n = 100000000
x, v, m = ic.sech2_disk(n, seed=2)
y = torch.linspace(-6, 6, 100001, dtype=torch.float64, device='cuda')
for _ in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    p = wendy_b200.potential(y, x, v, m)
    torch.cuda.synchronize()
    print('potential N=1e8 Y=1e5: %.1f ms' % (1e3 * (time.perf_counter() - t)))
    t = time.perf_counter()
    e = wendy_b200.energy(x, v, m, individual=True)
    torch.cuda.synchronize()
    print('individual energies N=1e8: %.1f ms' % (1e3 * (time.perf_counter() - t)))
print(float(p[50000]), float(e.sum()))

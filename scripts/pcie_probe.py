"""Host<->device copy rates on this box (what bounds the end-to-end generator path): D2H of 0.8 GB into
cudaHostAlloc'ed and into cudaHostRegister'ed memory, one stream and two concurrent streams; H2D likewise;
cost of cudaMalloc / cudaHostRegister of large blocks."""
import os
import sys
import time

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wendy_b200 import _lib

lib = _lib.load()
n = 100000000
torch.cuda.init()
t = time.perf_counter(); d = [torch.empty(n, dtype=torch.float64, device='cuda') for _ in range(2)]; torch.cuda.synchronize()
print('cudaMalloc 2 x 0.8 GB (torch)             %8.1f ms' % (1e3 * (time.perf_counter() - t)))
for x in d:
    x.normal_()
t = time.perf_counter(); hp = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(2)]
print('cudaHostAlloc 2 x 0.8 GB                  %8.1f ms' % (1e3 * (time.perf_counter() - t)))
hr = [numpy.empty(n) for _ in range(2)]
t = time.perf_counter()
for a in hr:
    assert lib.wendy_cuda_pin(a.ctypes.data, a.nbytes) == 0
print('cudaHostRegister 2 x 0.8 GB (untouched)   %8.1f ms' % (1e3 * (time.perf_counter() - t)))
hrt = [torch.from_numpy(a) for a in hr]
s = [torch.cuda.Stream() for _ in range(2)]


def timed(label, fn, nbytes, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    print('%-42s %8.1f ms  %6.1f GB/s' % (label, 1e3 * best, nbytes / best / 1e9))


def d2h_seq(h):
    with torch.cuda.stream(s[0]):
        h[0].copy_(d[0], non_blocking=True); h[1].copy_(d[1], non_blocking=True)


def d2h_par(h):
    for i in range(2):
        with torch.cuda.stream(s[i]):
            h[i].copy_(d[i], non_blocking=True)


def h2d_seq(h):
    with torch.cuda.stream(s[0]):
        d[0].copy_(h[0], non_blocking=True); d[1].copy_(h[1], non_blocking=True)


def h2d_par(h):
    for i in range(2):
        with torch.cuda.stream(s[i]):
            d[i].copy_(h[i], non_blocking=True)


timed('D2H 1.6 GB, cudaHostAlloc, one stream', lambda: d2h_seq(hp), 1.6e9)
timed('D2H 1.6 GB, cudaHostAlloc, two streams', lambda: d2h_par(hp), 1.6e9)
timed('D2H 1.6 GB, cudaHostRegister, one stream', lambda: d2h_seq(hrt), 1.6e9)
timed('D2H 1.6 GB, cudaHostRegister, two streams', lambda: d2h_par(hrt), 1.6e9)
timed('H2D 1.6 GB, cudaHostAlloc, one stream', lambda: h2d_seq(hp), 1.6e9)
timed('H2D 1.6 GB, cudaHostAlloc, two streams', lambda: h2d_par(hp), 1.6e9)
pg = [torch.from_numpy(numpy.random.rand(n)) for _ in range(2)]
timed('H2D 1.6 GB, pageable (driver staging)', lambda: h2d_seq(pg), 1.6e9, reps=2)

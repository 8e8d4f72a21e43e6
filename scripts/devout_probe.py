import sys, time
sys.path.insert(0, '/root/repo')
import numpy, torch
import wendy_b200
from bench import sech2_ic
n = 100000000
x, v, m = sech2_ic(n, 2)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    g = wendy_b200.nbody(x, v, m, 0.01, approx=True, nleap=10, omega=1.1, output='device')
    ts = []
    for i in range(6):
        xt, vt = next(g)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    g.close()
    print('rep', rep, ['%.3f' % t for t in ts], flush=True)

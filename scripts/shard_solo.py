"""One GPU, one rank: the sharded step kernel (persistent instance 3 + inject kernel) without any peer to wait for,
beside the plain instance on the same particles -- what the shard instance itself costs.
   WENDY_B200_SHARD_PEER=force python scripts/shard_solo.py [N]"""
import os
import sys
import time

os.environ['WENDY_B200_SHARD_PEER'] = 'force'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import torch

import wendy_b200
from wendy_b200 import multi
from bench import sech2_ic


class SoloComm(object):
    rank, size, device = 0, 1, 'cuda'

    def allgather_vec(self, vec):
        return numpy.asarray(vec, dtype=numpy.float64)[None, :].copy()

    def exchange(self, send, counts=None):
        return torch.zeros((0, 3), dtype=torch.float64, device='cuda')


n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
x, v, m = sech2_ic(n, 2)
s = multi.ShardedSystem(x, v, numpy.arange(n, dtype=numpy.int32), m[0], numpy.sum(m), SoloComm(), omega=1.1)
for _ in range(2):
    s.step(1e-3, 10)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
est = s.engine.stream
e0.record(est)
for _ in range(3):
    s.step(1e-3, 10)
e1.record(est)
est.synchronize()
print('shard instance (peer=%s): %.4f ms per sub-step' % (s.peer, e0.elapsed_time(e1) / 30), s.engine.stats(), flush=True)
s.close()
for cap, fill in ((0, 0), (2048, 1536)):
    st = wendy_b200.ApproxState(x, v, m, omega2=1.1 ** 2., cap=cap, fill=fill)
    for _ in range(2):
        st.step(1e-3, 10)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        st.step(1e-3, 10)
    e1.record()
    torch.cuda.synchronize()
    print('plain instance cap=%d fill=%d: %.4f ms per sub-step' % (cap, fill, e0.elapsed_time(e1) / 30), st.stats(), flush=True)
    st.close()

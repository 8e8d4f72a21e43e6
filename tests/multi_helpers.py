"""Test infrastructure for wendy_b200.multi: a numpy local engine (so the sharding logic runs
under gloo on CPUs) and an in-process communicator (ranks = threads, for single-GPU runs)."""
import threading

import numpy
import torch

from wendy_b200.multi import route


class NumpyShardEngine(object):
    """CPU stand-in for CudaShardEngine with identical semantics; arithmetic follows the oracle
    restatement (oracle/wendy_oracle.py::numpy_onestep, serial scan, equal masses)."""

    def __init__(self, x, v, ids, m0, totmass, omega2, nranks, rank, bounds, capacity, outbox_capacity):
        self.x, self.v, self.ids = numpy.array(x), numpy.array(v), numpy.array(ids, dtype=numpy.int32)
        self.m0, self.tot, self.omega2 = m0, totmass, omega2
        self.nranks, self.rank, self.bounds = nranks, rank, numpy.array(bounds)
        self.capacity = capacity

    def substep(self, h_pre, dt_kick, dt_drift, h_next, pc_offset):
        x = self.x + h_pre * self.v if h_pre != 0. else self.x
        order = numpy.lexsort((self.ids, x))
        xs, vs, ids = x[order], self.v[order], self.ids[order]
        # the reference's serial running sum (wendy.c:359-360) at the GLOBAL sorted position, as the CUDA engine
        cum = numpy.concatenate(([0.], numpy.cumsum(numpy.full(pc_offset + len(xs), self.m0))))[pc_offset:pc_offset + len(xs)]
        g = (self.tot - 2. * cum) - self.m0
        if self.omega2 >= 0:
            g = g - self.omega2 * xs
        v2 = vs + dt_kick * g
        x2 = xs + dt_drift * v2
        key = x2 + h_next * v2 if h_next != 0. else x2
        owner = route(key, self.bounds)
        keep = owner == self.rank
        out = [torch.from_numpy(numpy.stack((x2[owner == p], v2[owner == p], ids[owner == p].astype(numpy.float64)), axis=1))
               if p != self.rank else torch.zeros((0, 3), dtype=torch.float64) for p in range(self.nranks)]
        self.x, self.v, self.ids = x2[keep], v2[keep], ids[keep]
        return out

    def inject(self, packed):
        if packed.shape[0] == 0:
            return
        a = packed.cpu().numpy()
        self.x = numpy.concatenate((self.x, a[:, 0]))
        self.v = numpy.concatenate((self.v, a[:, 1]))
        self.ids = numpy.concatenate((self.ids, a[:, 2].astype(numpy.int32)))
        assert len(self.x) <= self.capacity

    def count(self):
        return len(self.x)

    def read(self):
        return self.ids, self.x, self.v

    def close(self):
        pass


class ThreadComm(object):
    """Communicator whose ranks are threads of one process (shared mailboxes + barriers)."""

    class _World(object):
        def __init__(self, size):
            self.size = size
            self.barrier = threading.Barrier(size)
            self.vec = [None] * size
            self.mail = [[None] * size for _ in range(size)]

    in_process = True  # the ranks share one process (and, on a GPU, one device)

    def __init__(self, world, rank, device='cpu'):
        self.w, self.rank, self.size, self.device = world, rank, world.size, device

    def barrier(self):
        self.w.barrier.wait()

    def allgather_vec(self, vec):
        self.w.vec[self.rank] = numpy.asarray(vec, dtype=numpy.float64).copy()
        self.w.barrier.wait()
        out = numpy.stack(self.w.vec)
        self.w.barrier.wait()
        return out

    def exchange(self, send, counts=None):
        for p in range(self.size):
            self.w.mail[p][self.rank] = send[p].clone()
        self.w.barrier.wait()
        recv = [self.w.mail[self.rank][p] for p in range(self.size) if p != self.rank]
        self.w.barrier.wait()
        return torch.cat(recv, dim=0) if recv else torch.zeros((0, 3), dtype=torch.float64)


def run_threads(size, fn, device='cpu'):
    """Run fn(comm) on `size` threads; re-raises the first exception; returns the results."""
    world = ThreadComm._World(size)
    res, err = [None] * size, []

    def work(r):
        try:
            res[r] = fn(ThreadComm(world, r, device))
        except BaseException as e:  # noqa: BLE001
            err.append(e)
            world.barrier.abort()
    ts = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if err:
        raise err[0]
    return res

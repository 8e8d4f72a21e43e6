"""CPU: the multi-GPU host logic of wendy_b200/multi.py (sample-sort partition, migrant exchange,
count offsets) with world_size 2 under gloo and with 3 in-process ranks, on a numpy local engine.
The sharded run must reproduce the single-process oracle bit for bit."""
import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_helpers import NumpyShardEngine, run_threads
from oracle import wendy_oracle as wo
from wendy_b200 import multi


def _problem(n=3000, seed=4):
    x, v, m = wo.sech2_ic(n, seed=seed)
    return x, v, m


def _oracle(x, v, m, dt_leap, nleap, calls, omega):
    om2 = -1. if omega is None else omega ** 2
    for _ in range(calls):
        x, v, _, _ = wo.numpy_onestep(x, v, m, numpy.sum(m), dt_leap, nleap, om2)
    return x, v


def _run_rank(comm, omega, dt_leap=0.01, nleap=4, calls=3, n=3000):
    x, v, m = _problem(n)
    mine = numpy.arange(n) % comm.size == comm.rank  # every rank starts with an arbitrary subset
    s = multi.ShardedSystem(x[mine], v[mine], numpy.arange(n)[mine], m[0], numpy.sum(m), comm, omega=omega,
                            engine_factory=NumpyShardEngine)
    for _ in range(calls):
        s.step(dt_leap, nleap)
    X, V = s.gather(n)
    xo, vo = _oracle(x, v, m, dt_leap, nleap, calls, omega)
    assert numpy.array_equal(X, xo) and numpy.array_equal(V, vo)
    assert abs(s.counts.sum() - n) == 0
    return s.migrated, s.counts


def _gloo_worker(rank, world, port, omega):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        migrated, counts = _run_rank(multi.TorchComm(device='cpu'), omega)
        assert migrated > 0  # the test must actually exercise the exchange
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('omega', [None, 1.1])
def test_sharded_system_world2_gloo_matches_oracle(omega):
    mp.spawn(_gloo_worker, args=(2, _free_port(), omega), nprocs=2, join=True)


def test_sharded_system_three_thread_ranks_matches_oracle():
    res = run_threads(3, lambda comm: _run_rank(comm, 0.7, dt_leap=0.02, nleap=5, calls=2))
    assert sum(r[0] for r in res) > 0
    assert abs(int(res[0][1].sum()) - 3000) == 0


def test_choose_bounds_and_route():
    b, share = multi.choose_bounds(numpy.linspace(0., 1., 1000), 4)
    assert numpy.allclose(share, 0.25, atol=2e-3)
    assert b[0] == -numpy.inf and b[-1] == numpy.inf and numpy.all(numpy.diff(b) >= 0)
    keys = numpy.array([-5., 0.1, 0.26, 0.6, 0.99, 7.])
    assert list(multi.route(keys, b)) == [0, 0, 1, 2, 3, 3]
    # a key equal to an edge belongs to the upper range (same convention as the bucket splitters)
    assert multi.route(numpy.array([b[2]]), b)[0] == 2


def test_cost_balanced_bounds_give_dense_ranges_fewer_particles():
    """Large N dt: a particle of the dense centre costs more (it crosses more buckets per sub-step), so the central
    ranks get fewer particles; without a displacement estimate the split is by count."""
    rs = numpy.random.RandomState(3)
    keys = numpy.arctanh(2. * rs.uniform(size=200000) - 1.) * 2.
    b0, s0 = multi.choose_bounds(keys, 8)
    assert numpy.allclose(s0, 0.125, atol=1e-3)
    b1, s1 = multi.choose_bounds(keys, 8, n_total=800000000, disp=1e-3)
    assert numpy.all(numpy.diff(b1) >= 0) and abs(s1.sum() - 1.) < 1e-12
    assert s1[3] < 0.115 and s1[4] < 0.115 and s1[0] > 0.135 and s1[7] > 0.135  # centre lighter, wings heavier
    # equal COST: weight 1 + 0.008 D_b per particle, D_b = density * disp / 1664
    ks = numpy.sort(keys)
    dens = 800000000 * 0.25 / numpy.cosh(0.5 * ks) ** 2
    w = 1. + 0.008 * dens * 1e-3 / 1664.
    cost = numpy.array([w[(ks >= b1[r]) & (ks < b1[r + 1])].sum() for r in range(8)])
    assert cost.max() / cost.min() < 1.05
    # small systems / small dt: nothing to balance
    b2, s2 = multi.choose_bounds(keys, 8, n_total=800000000, disp=1e-7)
    assert numpy.allclose(s2, 0.125, atol=2e-3)


def test_shard_ensemble_covers_everything_once():
    for n, w in ((4096, 8), (10, 3), (2, 4)):
        seen = [i for r in range(w) for i in multi.shard_ensemble(n, r, w)]
        assert seen == list(range(n))

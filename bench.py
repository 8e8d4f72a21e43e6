#!/usr/bin/env python
"""bench.py -- particle-steps/s of the approximate-integration hot path on B200.

Default workload = BASELINE.json configs[2], the configuration the metric is quoted on: sech^2 self-gravitating
disk with a harmonic term (omega = 1.1), N = 1e8 per GPU, fp64, equal masses, seeded numpy ICs.
A bench "step" is ONE call of the reference entry point (wendy/wendy.c:385-418) with nleap leapfrog sub-steps;
particle-steps = N * nleap * steps.

  python bench.py [--gpus N] [--steps K] [--warmup W]      our CUDA path; one JSON line (rank 0)
  python bench.py --impl reference ...                      the reference's own C path on the host cores
  python bench.py --config 1|2|3|4|5                        the other BASELINE configs as their own legs

The default line also carries: `variants` (config 3's other sub-runs: dt_leap 1e-5 and 5e-3, and the full radix
sort forced every sub-step), `parity` (x, v of the GPU against the compiled reference on the SAME N=1e8 system,
bit for bit) and `cpu_baseline` (the reference timed on that system).  N>1 is launched by torchrun, one rank per
GPU: the ranks hold ONE system of N*gpus particles, range-partitioned by position (BASELINE configs[3];
wendy_b200/multi.py, migrants exchanged over NVLink peer memory); `--mode ensemble` runs independent
realisations instead (configs[4]).  Per-GPU work is fixed either way: scaling is "weak".
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = 40.  # algorithmic bytes per particle-step: read x,v,m + write x,v
# dram__bytes_read.sum + dram__bytes_write.sum per particle of the dominant kernel, from the ncu --set full
# captures under profiles/ (persistent CTA kernel: profiles/r02/tile_1e8_dt1e-3_summary.txt; warp kernel r01)
TRAFFIC_B_PER_PARTICLE = {2048: 39.9, 256: 40.2}
KERNEL_NAME = {
    2: 'tile_kernel<2048,512,LOAD_BUCKET,EMIT_SPLITTER,PHYS,EQM,PERSIST=2> (persistent CTAs, two per SM; next bucket by TMA)',
    3: 'tile_kernel<2048,512,LOAD_BUCKET,EMIT_SPLITTER,PHYS,EQM,PERSIST=3> (same + migrants stored into the peers\' inboxes)',
    1: 'tile_kernel<2048,512,...,PERSIST=1> (persistent CTAs; ext-force / host-exchanged shard instance)',
    256: 'wstep_kernel<256,8,EQM> (one warp per bucket)',
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=3, choices=[1, 2, 3, 4, 5], help='BASELINE.json config (1-based)')
    ap.add_argument('--particles', dest='n', type=float, default=1e8, help='particles per GPU (config 3/4)')
    ap.add_argument('--leap', dest='nleap', type=int, default=10)
    ap.add_argument('--dt-leap', type=float, default=1e-3)
    ap.add_argument('--omega', type=float, default=1.1)
    ap.add_argument('--sort', default='gpu', choices=['gpu', 'gpu-radix'])
    ap.add_argument('--ref-particles', dest='ref_n', type=float, default=0, help='reference arm: particles (0 = same as ours)')
    ap.add_argument('--skip-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    ap.add_argument('--skip-e2e', dest='no_e2e', action='store_true')
    ap.add_argument('--skip-variants', dest='no_variants', action='store_true')
    ap.add_argument('--cap', type=int, default=0, help='bucket capacity: 0 = library choice, 256 = warp kernel, 2048 = CTA kernel')
    ap.add_argument('--fill', type=int, default=0, help='target particles per bucket (0 = library default)')
    ap.add_argument('--mode', default='auto', choices=['auto', 'ensemble', 'sharded'],
                    help='N>1: independent realisations per GPU, or ONE system of n*gpus particles '
                         'range-partitioned over the GPUs (auto = sharded)')
    return ap.parse_args()


def sech2_ic(n, seed):
    """reference examples/WendyScaling.ipynb:57-65."""
    rs = numpy.random.RandomState(seed)
    x = numpy.arctanh(2. * rs.uniform(size=n) - 1.) * 2.
    v = rs.normal(size=n)
    v -= numpy.mean(v)
    m = numpy.full(n, 1. / n)
    return x, v, m


def slab_ic(n, seed=3):
    """config 2 (survey-defined cold slab: the violent-relaxation notebook is missing from the snapshot)."""
    rs = numpy.random.RandomState(seed)
    x = rs.uniform(-0.5, 0.5, size=n)
    v = 0.05 * rs.normal(size=n)
    v -= numpy.mean(v)
    return x, v, numpy.full(n, 1. / n)


def host_info():
    gcc = ''
    try:
        gcc = subprocess.run(['gcc', '--version'], capture_output=True, text=True, timeout=5).stdout.splitlines()[0]
    except Exception:
        pass
    return {'nproc': os.cpu_count(), 'OMP_NUM_THREADS': os.environ.get('OMP_NUM_THREADS'), 'gcc': gcc,
            'PARALLEL_SORT_NUM_THREADS': '32 (wendy_c.so, the reference default, parallel_sort.h:7-14) / nproc (wendy_c_nt.so)'}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': float(self.rows[0][1]) if self.rows else None,
                'reasons': reasons, 'samples': len(self.rows)}


def make_reference(x, v, m, omega, dt_leap, sort='parallel', variant='wendy_c.so', nleap=1):
    """The reference's own C path (oracle/_ref/*.so = /root/reference/wendy/*.c compiled unmodified), driven as
    wendy/wendy.py:425-433 drives it; falls back to our C restatement when the reference could not be built."""
    from oracle import wendy_oracle as wo
    if wo.reference_available(variant):
        return wo.Reference(x, v, m, dt_leap * nleap, nleap, omega=omega, sort=sort, variant=variant), 'reference'
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'oracle'])
    return wo.COracle(x, v, m, dt_leap * nleap, nleap, omega=omega), 'port'


def time_reference(r, steps, warmup):
    for _ in range(warmup):
        r.step()
    ts = []
    for _ in range(steps):
        t = time.perf_counter()
        r.step()
        ts.append(time.perf_counter() - t)
    return sum(ts) / max(1, len(ts)), ts


def relerr(a, b):
    floor = 0.1 * numpy.sqrt(numpy.mean(b ** 2)) + 1e-300
    return float(numpy.max(numpy.abs(a - b) / numpy.maximum(floor, numpy.abs(b))))


def reference_arm(a, rank):
    """bench.py --impl reference: the reference C path on the host cores, SAME system as our arm (config 3:
    N = 1e8), each step a bounded sample of the workload: one leapfrog sub-step (a call with nleap = 1)."""
    if rank != 0:
        return
    cores = os.cpu_count()
    n = int(a.ref_n) if a.ref_n else int(a.n)
    if a.config == 2:
        x, v, m = slab_ic(1000000)
        omega, dt_leap, label = None, 0.005, 'cold slab N=1e6 (config 2), dt_leap=0.005'
    elif a.config == 1:
        x, v, m = sech2_ic(10000, 2)
        omega, dt_leap, label = None, 0.05, 'sech2 disk N=1e4 (config 1), dt=0.05, nleap=1'
    else:
        x, v, m = sech2_ic(n, 2)
        omega, dt_leap, label = a.omega, a.dt_leap, 'sech2 disk + harmonic omega=%g, N=%d, dt_leap=%g' % (a.omega, n, a.dt_leap)
    r, kind = make_reference(x, v, m, omega, dt_leap)
    per, _ = time_reference(r, a.steps, a.warmup)
    val = len(x) / per
    extra = {}
    try:  # the nproc-thread build of the parallel sort and the best serial sort, for context
        nb = min(len(x), 10000000)
        for name, sort, variant in (('parallel_sort_nproc_threads', 'parallel', 'wendy_c_nt.so'),
                                    ('serial_tim', 'tim', 'wendy_c.so'), ('serial_quick', 'quick', 'wendy_c.so')):
            rr, k2 = make_reference(x[:nb], v[:nb], m[:nb] * (len(x) / nb), omega, dt_leap, sort=sort, variant=variant)
            if k2 != 'reference':
                continue
            p2, _ = time_reference(rr, 2, 1)
            extra[name] = {'value': nb / p2, 'N': nb}
    except Exception as exc:  # noqa: BLE001
        extra['error'] = str(exc)[:200]
    sample = ('the whole N=%d system, one leapfrog sub-step (nleap=1 call) per step, sort=parallel' % len(x))
    print(json.dumps({
        'impl': 'reference', 'metric': 'particle-steps/s', 'value': val, 'unit': 'particle-steps/s',
        'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': per * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': label, 'host': host_info()},
        'cpu_baseline': {'value': val, 'unit': 'particle-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'other_reference_sorts': extra,
        'e2e': {'value': val, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    a = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    n = int(a.n)
    cores = os.cpu_count()
    # both arms may use every host thread (reference: its OpenMP loops; ours: input validation, staging of
    # pageable uploads and the bounce-buffered read-out); libgomp reads this when the libraries are loaded
    os.environ.pop('OMP_NUM_THREADS', None)

    if a.impl == 'reference':
        reference_arm(a, rank)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import wendy_b200
    if world > 1:
        # one rank per GPU: the library's host-side copy threads share the node's cores between the ranks
        from wendy_b200 import _lib as _wl
        _wl.load().wendy_host_set_threads(max(1, min(32, cores // int(os.environ.get('LOCAL_WORLD_SIZE', world)) - 1)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(t):
        if world == 1:
            return t
        tt = torch.tensor([t], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6650.)
    stream = torch.cuda.current_stream().cuda_stream
    stat_keys = ('substeps', 'rebuilds', 'failed_substeps', 'kernel_launches', 'radix_fallbacks', 'left_window')

    def stat_delta(s1, s0):
        d = {k: s1[k] - s0[k] for k in stat_keys}
        d['max_bucket_count'] = s1['max_bucket_count']
        d['cap'] = s1['cap']
        d['buckets'] = s1['buckets']
        return d

    def run_single(x, v, m, omega2, dt_leap, sort, steps, warmup, nleap, n_segments=1, cap=0, fill=0):
        st = wendy_b200.ApproxState(x, v, m, omega2=omega2, sort=sort, stream=stream, cap=cap, fill=fill,
                                    n_segments=n_segments)
        for _ in range(warmup):
            st.step(dt_leap, nleap)
        s0 = st.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            st.step(dt_leap, nleap)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        d = stat_delta(st.stats(), s0)
        st.close()
        return ms, d

    # ------------------------------------------------------------------------------------------------------
    # the other BASELINE configs as their own legs
    if a.config in (1, 2, 5):
        out = None
        if a.config == 1:  # N=1e4, dt=0.05, 100 outputs: the reference's own CPU-runnable case (KAT-D)
            x, v, m = sech2_ic(10000, 2)
            g = wendy_b200.nbody(x, v, m, 0.05, approx=True, nleap=1)
            next(g)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(99):
                xo, vo = next(g)
            el = time.perf_counter() - t0
            hx, hv = hashlib.sha256(xo.tobytes()).hexdigest()[:16], hashlib.sha256(vo.tobytes()).hexdigest()[:16]
            g.close()
            out = {'value': 1e4 * 99 / el, 'ms_per_step': 1e3 * el / 99,
                   'config': {'workload': 'config 1: sech2 disk N=1e4, dt=0.05, nleap=1, 100 outputs through nbody()'},
                   'parity': {'sha256_x': hx, 'sha256_v': hv,
                              'equals_reference_KAT_D': hx == '21614814a2a156b6' and hv == '7730018ad0084bb4'}}
        if a.config == 2:  # cold slab N=1e6, 1000 leapfrog steps, energy-drift statistics beside the reference's own
            from oracle import wendy_oracle as wo
            nn, nsteps, dtl, nl = 1000000, 1000, 0.005, 10  # 100 calls of nleap = 10 (the host syncs once per call)
            x, v, m = slab_ic(nn)
            E0 = wo.energy(x, v, m)
            st = wendy_b200.ApproxState(x, v, m, stream=stream)
            st.step(dtl, nl)
            s0 = st.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            drift = []
            torch.cuda.synchronize()
            e0.record()
            for i in range(1, nsteps // nl):
                st.step(dtl, nl)
                if (i + 1) % 10 == 0:
                    ke, he, pe, _ = st.energy_terms()
                    drift.append(abs((ke + he + pe) - E0) / abs(E0))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            d = stat_delta(st.stats(), s0)
            st.close()

            def timed(nleap_v, calls, **kw):
                stv = wendy_b200.ApproxState(x, v, m, stream=stream, **kw)
                stv.step(dtl, nleap_v)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(calls):
                    stv.step(dtl, nleap_v)
                e1.record()
                torch.cuda.synchronize()
                r_ = float(nn) * nleap_v * calls / (e0.elapsed_time(e1) * 1e-3)
                stv.close()
                return r_
            var = {'300 calls of nleap=1 (a host round trip per leapfrog step)': {'value': timed(1, 300)},
                   'sort=gpu-radix, 30 calls of nleap=10 (full sort every step, no layout to rebuild)':
                       {'value': timed(nl, 30, sort='gpu-radix')}}
            ref_drift, ref_rate, same10 = None, None, None
            if rank == 0 and not a.no_cpu_baseline:
                r, kind = make_reference(x, v, m, None, dtl, nleap=nl)
                t0 = time.perf_counter()
                ref_drift = []
                for i in range(nsteps // nl):
                    xr, vr = r.step()
                    if i == 0:
                        g10 = wendy_b200.ApproxState(x, v, m, stream=stream)
                        g10.step(dtl, nl)
                        x10, v10 = g10.read()
                        g10.close()
                        same10 = bool(numpy.array_equal(x10, xr) and numpy.array_equal(v10, vr))
                    if (i + 1) % 10 == 0:
                        ref_drift.append(abs(wo.energy(xr, vr, m) - E0) / abs(E0))
                ref_rate = nn * nsteps / (time.perf_counter() - t0)
            out = {'value': float(nn) * (nsteps - nl) / (ms * 1e-3), 'ms_per_step': ms / (nsteps - nl),
                   'config': {'workload': 'config 2: cold slab N=1e6, omega=None, dt_leap=0.005, 1000 leapfrog steps as 100 calls '
                                          'of nleap=10 (violent relaxation)'},
                   'path_stats': d, 'gpu_launches': d['kernel_launches'], 'variants': var,
                   'energy_drift': {'gpu_abs_dE_over_E_every_100_steps': drift, 'reference_same_ICs': ref_drift,
                                    'note': 'chaotic after a few steps: compared as statistics, not particle by particle'},
                   'parity': {'bit_identical_to_reference_after_10_steps': same10},
                   'cpu_baseline': {'value': ref_rate, 'unit': 'particle-steps/s', 'cores': cores, 'kind': 'reference',
                                    'sample': 'the same 1000 steps (100 calls of nleap=10), sort=parallel'}}
        if a.config == 5:  # Gaia phase-space spiral ensemble: realisations of 1e5 particles + torch ext_force
            from wendy_b200 import multi
            total_real = 4096 if world == 8 else 512 * world
            mine = multi.shard_ensemble(total_real, rank, world)
            S, L = len(mine), 100000
            alpha, sigma, zh = 0.3, 1., 1.
            xs, vs = numpy.empty(S * L), numpy.empty(S * L)
            for j, r_ in enumerate(mine):  # per realisation RandomState(2 + r), SURVEY 8d config 5
                rs = numpy.random.RandomState(2 + r_)
                xs[j * L:(j + 1) * L] = numpy.arctanh(2. * rs.uniform(size=L) - 1.) * 2. * zh
                vv = rs.normal(size=L) * sigma
                vs[j * L:(j + 1) * L] = vv - numpy.mean(vv) + sigma  # the kick that winds up the spiral
            ms_ = numpy.full(S * L, alpha / L)
            F = lambda xx, t: -(1. - alpha) * sigma ** 2. * torch.tanh(0.5 * xx / zh) / zh  # noqa: E731
            barrier()
            t_c = time.perf_counter()
            g = wendy_b200.nbody(xs, vs, ms_, 0.05, approx=True, nleap=10, ext_force=F, n_segments=S)
            next(g)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                next(g)
            torch.cuda.synchronize()
            el = max_over_ranks(time.perf_counter() - t0)
            g.close()
            out = {'value': float(total_real) * L * 10 * a.steps / el, 'ms_per_step': 1e3 * el / a.steps,
                   'config': {'workload': 'config 5: Gaia phase-space-spiral ensemble, %d realisations x 1e5 particles (%d per GPU), '
                                          'torch ext_force, dt=0.05, nleap=10, through nbody() incl. D2H of x, v every output'
                                          % (total_real, S), 'parallelism': 'realisations dealt out to the ranks, no collective'},
                   'construction_s': t0 - t_c}
        if rank == 0:
            base = {'metric': 'particle-steps/s', 'unit': 'particle-steps/s', 'n_gpus': world, 'steps': a.steps,
                    'warmup': a.warmup, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                    'dtype': 'f64', 'data': 'synthetic'}
            base.update(out)
            print(json.dumps(base))
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------------------------------------------
    # config 3 (N=1) / config 4 (N>1): the headline line
    x, v, m = sech2_ic(n, 2 + rank)
    omega2 = a.omega ** 2

    def run_sharded(dt_leap, steps, warmup, nleap):
        """ONE system of n*world particles, range-partitioned by position over the ranks."""
        from wendy_b200 import multi
        comm = multi.TorchComm(device='cuda')
        ids = numpy.arange(n, dtype=numpy.int64) + rank * n
        m0 = 1. / (n * world)
        s = multi.ShardedSystem(x, v, ids.astype(numpy.int32), m0, m0 * n * world, comm, omega=a.omega)
        for _ in range(warmup):
            s.step(dt_leap, nleap)
        est = s.engine.stream  # the shard's kernels run on the engine's own stream
        s0 = s.engine.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mig0 = s.migrated
        tm0 = dict(s.timing)
        barrier()
        e0.record(est)
        for _ in range(steps):
            s.step(dt_leap, nleap)
        e1.record(est)
        barrier()
        est.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1))
        d = stat_delta(s.engine.stats(), s0)
        d['exchange'] = 'device-driven over peer memory (NVLink)' if s.peer else 'host-orchestrated (NCCL all-gather + all-to-all per sub-step)'
        d['migrants_per_substep_rank0'] = (s.migrated - mig0) / float(steps * nleap)
        d['particles_per_rank'] = [int(c) for c in s.counts]
        d['host_ms_per_call_rank0'] = {k: 1e3 * (t - tm0[k]) / max(1, steps) for k, t in s.timing.items()}
        peer = s.peer
        s.close()
        return ms, d, peer

    def sharded_check():
        """Correctness inside the scaling run: a 2.4e6-particle system stepped sharded over all ranks must equal the
        same system stepped on one GPU (rank 0), bit for bit (sha256 of x and v in particle order)."""
        from wendy_b200 import multi
        nn = 2400000
        xx, vv, mm = sech2_ic(nn, 6)
        comm = multi.TorchComm(device='cuda')
        mine = numpy.arange(nn) % world == rank
        s = multi.ShardedSystem(xx[mine], vv[mine], numpy.arange(nn, dtype=numpy.int32)[mine], mm[0], numpy.sum(mm),
                                comm, omega=a.omega)
        s.step(1e-3, 4)
        s.step(1e-3, 4)
        X, V = s.gather(nn)
        s.close()
        res = None
        if rank == 0:
            st = wendy_b200.ApproxState(xx, vv, mm, omega2=omega2)
            st.step(1e-3, 4)
            st.step(1e-3, 4)
            Xs, Vs = st.read()
            st.close()
            h1 = hashlib.sha256(X.tobytes() + V.tobytes()).hexdigest()[:16]
            h2 = hashlib.sha256(Xs.tobytes() + Vs.tobytes()).hexdigest()[:16]
            res = {'particles': nn, 'substeps': 8, 'ranks': world, 'sha256_sharded': h1, 'sha256_one_gpu': h2,
                   'bit_identical': h1 == h2}
        return res

    def e2e_sharded(dt_leap, steps_e, nleap):
        """Public multi-GPU API with host buffers: construction (key sample, all-to-all partition, H2D) +
        steps, each followed by read_local() (compaction + D2H of the particles this rank owns)."""
        from wendy_b200 import multi
        barrier()
        t0 = time.perf_counter()
        comm = multi.TorchComm(device='cuda')
        ids = (numpy.arange(n, dtype=numpy.int64) + rank * n).astype(numpy.int32)
        m0 = 1. / (n * world)
        s = multi.ShardedSystem(x, v, ids, m0, m0 * n * world, comm, omega=a.omega)
        got = 0
        t1 = time.perf_counter()
        s.step(dt_leap, nleap)  # (the first call also partitions: sample sort + all-to-all + H2D + layout build)
        t2 = time.perf_counter()
        s.read_local_begin()
        for _ in range(steps_e - 1):
            # the read-out of output k (compaction, then D2H on a copy stream) runs beside the kernels of call k+1
            s.step(dt_leap, nleap)
            il, xl, vl = s.read_local_end()
            got = len(il)
            s.read_local_begin()
        il, xl, vl = s.read_local_end()
        got = len(il)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        el = max_over_ranks(t3 - t0)
        s.close()
        return {'value': float(n) * world * nleap * steps_e / el, 'unit': 'particle-steps/s',
                'h2d_bytes_per_step': 20. * n / steps_e, 'd2h_bytes_per_step': 20. * got,
                'phases_s_rank0': {'construct': t1 - t0, 'first_call_incl_partition': t2 - t1,
                                   'other_calls_and_readouts': t3 - t2},
                'note': 'multi.ShardedSystem from host arrays (sample-sort partition + H2D, amortised over %d steps) '
                        '+ step() + read-out (D2H of x, v, id of the owned particles) each step, the read-out of '
                        'output k overlapped with call k+1' % steps_e}

    sharded = world > 1 and a.mode in ('auto', 'sharded')
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    peer = False
    if sharded:
        ms, d, peer = run_sharded(a.dt_leap, a.steps, a.warmup, a.nleap)
    else:
        ms, d = run_single(x, v, m, omega2, a.dt_leap, a.sort, a.steps, a.warmup, a.nleap, cap=a.cap, fill=a.fill)
    clocks = sampler.stop() if sampler else None
    psteps = float(n) * world * a.nleap * a.steps
    value = psteps / (ms * 1e-3)

    # N>1, sharded: also time the other multi-GPU mode of the north star -- independent realisations, one per
    # GPU, no data-path collective (BASELINE.json configs[4]) -- so that both scalings can be read off one
    # line.  No collective inside the leg (a rank that fails must not hang the others): local events, then one
    # max over ranks.
    ensemble, check, regime = None, None, None
    if sharded and not a.no_variants:
        # The cost of a sub-step grows with the displacement per sub-step measured in ranks, N_total * rho * v * dt:
        # at a fixed dt a system of `world` times more particles sends every particle across `world` times more
        # buckets (path_stats.left_window: particles leaving the destination window, which the library widens from 256 to at most 1024 buckets on that evidence).  The same run
        # with dt_leap / world keeps that displacement -- the regime of the N=1 line -- and isolates what the
        # exchange itself costs.
        try:
            vms, vd, _ = run_sharded(a.dt_leap / world, 3, 2, a.nleap)
            regime = {'dt_leap': a.dt_leap / world, 'value': float(n) * world * a.nleap * 3 / (vms * 1e-3),
                      'ms_per_substep': vms / (3 * a.nleap), 'left_window': vd['left_window'],
                      'note': 'same system, dt_leap / n_gpus: displacement per sub-step in ranks as in the N=1 run'}
        except Exception as exc:  # noqa: BLE001
            regime = {'error': str(exc)[:200]}
    if sharded:
        steps_v = max(2, a.steps // 2)
        ms_loc = -1.
        try:
            st = wendy_b200.ApproxState(x, v, m, omega2=omega2, stream=stream)
            for _ in range(2):
                st.step(a.dt_leap, a.nleap)
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            v0.record()
            for _ in range(steps_v):
                st.step(a.dt_leap, a.nleap)
            v1.record()
            torch.cuda.synchronize()
            ms_loc = v0.elapsed_time(v1)
            st.close()
        except Exception as exc:  # noqa: BLE001 -- reported below, the main number stands
            ensemble = {'error': str(exc)[:200]}
        failed_any = max_over_ranks(1. if ms_loc < 0 else 0.) > 0.
        ms_max = max_over_ranks(ms_loc)
        if ensemble is None and not failed_any:
            ensemble = {'value': float(n) * world * a.nleap * steps_v / (ms_max * 1e-3), 'unit': 'particle-steps/s',
                        'note': 'independent realisations of %d particles, one per GPU, no data-path collective; '
                                '%d timed calls, max over ranks of local CUDA-event times' % (n, steps_v)}
        elif ensemble is None:
            ensemble = {'error': 'another rank failed'}
        try:
            check = sharded_check()
        except Exception as exc:  # noqa: BLE001
            check = {'error': str(exc)[:300]}

    # dominant kernel: the step kernel, one launch per sub-step on the bucket path.  Its average duration is the
    # timed region / sub-steps (CUDA events on the launching stream); the count-prefix kernel that precedes each
    # launch (and, sharded, the inject kernel that follows it) are inside that figure (0.3 % by the ncu launch list)
    launches_step = max(1, d['substeps'])
    ms_per_launch = ms / launches_step
    achieved = ALG_BYTES * n / (ms_per_launch * 1e-3) / 1e9
    if a.sort != 'gpu':
        kname = 'radix passes (onesweep) + tile_kernel<LOAD_GATHER,EMIT_RANK>'
    elif d['cap'] == 256:
        kname = KERNEL_NAME[256]
    else:
        kname = KERNEL_NAME[3 if (sharded and peer) else (1 if sharded else 2)]
    traffic = TRAFFIC_B_PER_PARTICLE.get(d['cap'], None)
    out = {
        'metric': 'particle-steps/s', 'value': value, 'unit': 'particle-steps/s', 'n_gpus': world,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'sech2 disk + harmonic omega=%g, N=%d per GPU, dt_leap=%g, nleap=%d, sort=%s, equal masses'
                               % (a.omega, n, a.dt_leap, a.nleap, a.sort),
                   'parallelism': ('one system of %d particles range-partitioned over %d GPUs (sample-sort partition; '
                                   'per sub-step: %s)' % (n * world, world, d.get('exchange', ''))) if sharded
                   else ('independent realisations, one per GPU' if world > 1 else 'single GPU'),
                   'l2': 'state (%.1f GB) is far larger than L2' % (n * 28 / 1e9)},
        'gpu_launches': d['kernel_launches'],
        'path_stats': d,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                     'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback', 'unit': 'GB/s',
                     'frac': achieved / peak,
                     'traffic': traffic * n if (traffic and a.sort == 'gpu' and world == 1) else None,
                     'kernel': kname, 'ms_per_launch': ms_per_launch},
        'clocks': clocks,
    }
    if ensemble is not None:
        out['ensemble_mode'] = ensemble
    if check is not None:
        out['sharded_check'] = check
    if regime is not None:
        out['variants'] = {'dt_leap/n_gpus': regime}

    # ---- config 3's other sub-runs: dt_leap 1e-5 / 5e-3 and the full radix sort forced every sub-step -------
    if world == 1 and not a.no_variants:
        var = {}
        for dtl, srt, nl, stp in ((1e-5, 'gpu', a.nleap, 3), (5e-3, 'gpu', a.nleap, 3), (a.dt_leap, 'gpu-radix', 2, 2)):
            try:
                vms, vd = run_single(x, v, m, omega2, dtl, srt, stp, 2, nl)
                var['dt_leap=%g,%s' % (dtl, srt)] = {'value': float(n) * nl * stp / (vms * 1e-3), 'ms_per_substep': vms / (nl * stp),
                                                     'stats': vd}
            except Exception as exc:  # noqa: BLE001
                var['dt_leap=%g,%s' % (dtl, srt)] = {'error': str(exc)[:200]}
        # unequal masses (the reference's force takes arbitrary m): the general path -- exact 128-bit mass scan
        try:
            mg = m * (1. + 0.3 * (2. * numpy.random.RandomState(7).uniform(size=n) - 1.))
            vms, vd = run_single(x, v, mg, omega2, a.dt_leap, 'gpu', 3, 2, a.nleap)
            var['unequal masses,dt_leap=%g' % a.dt_leap] = {'value': float(n) * a.nleap * 3 / (vms * 1e-3),
                                                            'ms_per_substep': vms / (a.nleap * 3), 'stats': vd}
            del mg
        except Exception as exc:  # noqa: BLE001
            var['unequal masses'] = {'error': str(exc)[:200]}
        # GPU yard-stick for the radix path: torch.sort of N fp64 keys on this GPU -- CUB's
        # DeviceRadixSort on 64-bit keys with 64-bit indices -- beside the library's own (u64 key, u32 index)
        # onesweep sort, which the forced-radix leg above runs once per sub-step (8.8 ms at N=1e8 on its own,
        # profiles/r02/kernels_tour_N1e8.md).  Library code: a reference point, not on any product path.
        try:
            keys = torch.randn(n, dtype=torch.float64, device='cuda')
            torch.sort(keys)
            y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            y0.record()
            for _ in range(3):
                torch.sort(keys)
            y1.record()
            torch.cuda.synchronize()
            out['yardstick'] = {'what': 'torch.sort of N fp64 keys on this GPU (CUB radix sort; library code, a reference point only)',
                                'ms_per_sort': y0.elapsed_time(y1) / 3.}
            del keys
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            out['yardstick'] = {'error': str(exc)[:200]}
        out['variants'] = var

    # ---- end to end through the public generator API, host buffers --------------------------
    if not a.no_e2e and sharded:
        out['e2e'] = e2e_sharded(a.dt_leap, max(2, min(a.steps, 20)), a.nleap)
    if not a.no_e2e and not sharded:
        steps_e = max(2, min(a.steps, 40))  # the K steps of the contract (bounded: each yields 1.6 GB to the host)
        barrier()
        t0 = time.perf_counter()
        g = wendy_b200.nbody(x, v, m, a.dt_leap * a.nleap, approx=True, nleap=a.nleap, omega=a.omega, sort=a.sort)
        for _ in range(steps_e):
            xo, vo = next(g)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        g.close()
        el = max_over_ranks(el)
        out['e2e'] = {'value': float(n) * world * a.nleap * steps_e / el, 'unit': 'particle-steps/s',
                      'h2d_bytes_per_step': 24. * n / steps_e, 'd2h_bytes_per_step': 16. * n,
                      'note': 'wendy_b200.nbody(): generator construction (H2D of x,v,m, amortised over %d steps) '
                              '+ next() x %d, each with nleap=%d sub-steps and a D2H of x,v' % (steps_e, steps_e, a.nleap)}

    # ---- the reference on the SAME system: timing (cpu_baseline) and parity at the headline size ----------
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r, kind = make_reference(x, v, m, a.omega, a.dt_leap)
        nsub = 4  # 1 warm-up (the first sort starts from random order) + 3 timed sub-steps
        # Exactly coincident keys: the reference's sort='parallel' (OpenMP merge over qsort leaves,
        # wendy/parallel_sort.c:95-137) leaves their order unspecified, ours is (x, index).  Two tied particles whose
        # ranks are swapped differ by 2 m0 dt in v -- not an error of either side; they are counted and excluded.
        def tied_now(xc, vc):
            key = xc + (a.dt_leap / 2.) * vc  # the key of the call's only force evaluation (wendy/wendy.c:398)
            perm = wendy_b200.argsort(key)
            ks = key[perm]
            eq = numpy.flatnonzero(ks[1:] == ks[:-1])
            return numpy.concatenate((perm[eq], perm[eq + 1]))
        tied = numpy.zeros(n, dtype=bool)
        g = wendy_b200.nbody(x, v, m, a.dt_leap, approx=True, nleap=1, omega=a.omega)
        ts, par = [], {}
        xc, vc = x, v
        for i in range(nsub):
            tied[tied_now(xc, vc)] = True
            t0 = time.perf_counter()
            xr, vr = r.step()
            ts.append(time.perf_counter() - t0)
            xc, vc = next(g)
            if i in (0, nsub - 1):
                bad = (xc != xr) | (vc != vr)
                par['after_%d_substeps' % (i + 1)] = {
                    'x_bit_identical': bool(numpy.array_equal(xc[~tied], xr[~tied])),
                    'v_bit_identical': bool(numpy.array_equal(vc[~tied], vr[~tied])),
                    'particles_in_exact_key_ties_so_far': int(tied.sum()),
                    'particles_differing': int(bad.sum()), 'of_which_tied': int((bad & tied).sum()),
                    'max_rel_err_x': relerr(xc, xr), 'max_rel_err_v': relerr(vc, vr)}
        g.close()
        per = sum(ts[1:]) / (nsub - 1)
        out['cpu_baseline'] = {'value': n / per, 'unit': 'particle-steps/s', 'cores': cores, 'kind': kind,
                               'sample': 'the same N=%d system: 1 warm-up + 3 timed leapfrog sub-steps (nleap=1 calls), reference '
                                         'sort=parallel, %.2f s per sub-step' % (n, per), 'host': host_info()}
        par['against'] = 'oracle/_ref/wendy_c.so (the unmodified reference C path), same ICs, N=%d, dt_leap=%g, nleap=1 calls' % (n, a.dt_leap)
        par['tolerance_north_star'] = ('x, v relative 1e-12 after 1 step, 1e-9 after 10; required here: bit-identical for every particle '
                                       'outside exact key ties (x_bit_identical / v_bit_identical are evaluated on those)')
        out['parity'] = par
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""bench.py -- particle-steps/s of the approximate-integration hot path on B200.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): sech^2
self-gravitating disk with a harmonic term (omega=1.1), N=1e8 per GPU, fp64, seeded numpy ICs
(SURVEY.md section 8d).  A bench "step" is ONE call of the reference entry point
(wendy/wendy.c:385-418) with nleap leapfrog sub-steps; particle-steps = N * nleap * steps.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference ...                          the reference's own C path on the host

N>1 is launched by torchrun (one rank per GPU).  By default the ranks then hold ONE system of N*gpus
particles, range-partitioned by position (wendy_b200/multi.py: all-to-all of migrants + all-gather of
counts per sub-step; BASELINE.json configs[3]); --mode ensemble runs independent realisations instead
(configs[4], no data-path collective).  Per-GPU work is fixed either way, so scaling is "weak".
"""
import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--particles', dest='n', type=float, default=1e8, help='particles per GPU')
    ap.add_argument('--leap', dest='nleap', type=int, default=10)
    ap.add_argument('--dt-leap', type=float, default=1e-3)
    ap.add_argument('--omega', type=float, default=1.1)
    ap.add_argument('--sort', default='gpu', choices=['gpu', 'gpu-radix'])
    ap.add_argument('--ref-particles', dest='ref_n', type=float, default=1e7, help='particles in the CPU sample')
    ap.add_argument('--skip-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    ap.add_argument('--skip-e2e', dest='no_e2e', action='store_true')
    ap.add_argument('--cap', type=int, default=0, help='bucket capacity: 0/256 = warp kernel, 2048 = CTA kernel')
    ap.add_argument('--fill', type=int, default=0, help='target particles per bucket (0 = library default)')
    ap.add_argument('--variants', action='store_true', help='also time other dt_leap / sort settings')
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


def time_reference(x, v, m, omega, dt_leap, steps, warmup):
    """The reference's own C path (oracle/_ref/wendy_c.so, sort='parallel', all host threads)."""
    from oracle import wendy_oracle as wo
    kind = 'reference' if wo.reference_available() else 'port'
    if kind == 'reference':
        r = wo.Reference(x, v, m, dt_leap, 1, omega=omega, sort='parallel')
    else:
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'oracle'])
        r = wo.COracle(x, v, m, dt_leap, 1, omega=omega)
    for _ in range(warmup):
        r.step()
    t = time.perf_counter()
    for _ in range(steps):
        r.step()
    el = time.perf_counter() - t
    return len(x) * steps / el, el / steps, kind


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
        if rank != 0:
            return
        os.environ.pop('OMP_NUM_THREADS', None)
        nr = int(a.ref_n)
        x, v, m = sech2_ic(nr, 2)
        val, per, kind = time_reference(x, v, m, a.omega, a.dt_leap, a.steps, a.warmup)
        sample = ('N=%d of the N=%d workload, one sub-step per step, sort=parallel '
                  '(PARALLEL_SORT_NUM_THREADS=32 default), OMP threads=%d' % (nr, n, cores))
        print(json.dumps({
            'impl': 'reference', 'metric': 'particle-steps/s', 'value': val, 'unit': 'particle-steps/s',
            'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': per * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': 'sech2 disk + harmonic omega=%g, dt_leap=%g (CPU sample N=%d)' % (a.omega, a.dt_leap, nr)},
            'cpu_baseline': {'value': val, 'unit': 'particle-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': val, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import wendy_b200

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

    x, v, m = sech2_ic(n, 2 + rank)
    omega2 = a.omega ** 2
    stream = torch.cuda.current_stream().cuda_stream

    def run(dt_leap, sort, steps, warmup, nleap):
        st = wendy_b200.ApproxState(x, v, m, omega2=omega2, sort=sort, stream=stream, cap=a.cap, fill=a.fill)
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
        s1 = st.stats()
        st.close()
        run.cap = s1['cap']
        d = {k: s1[k] - s0[k] for k in ('substeps', 'rebuilds', 'failed_substeps', 'kernel_launches', 'radix_fallbacks')}
        d['max_bucket_count'] = s1['max_bucket_count']
        d['left_window'] = s1['left_window'] - s0['left_window']
        return ms, d

    def run_sharded(dt_leap, steps, warmup, nleap):
        """ONE system of n*world particles, range-partitioned by position over the ranks."""
        from wendy_b200 import multi
        comm = multi.TorchComm(device='cuda')
        ids = numpy.arange(n, dtype=numpy.int64) + rank * n
        m0 = 1. / (n * world)
        s = multi.ShardedSystem(x, v, ids.astype(numpy.int32), m0, m0 * n * world, comm, omega=a.omega)
        for _ in range(warmup):
            s.step(dt_leap, nleap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mig0 = s.migrated
        barrier()
        e0.record()
        for _ in range(steps):
            s.step(dt_leap, nleap)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        d = {'substeps': steps * nleap, 'rebuilds': 0, 'failed_substeps': 0, 'kernel_launches': 2 * steps * nleap,
             'radix_fallbacks': 0, 'max_bucket_count': 0, 'left_window': 0,
             'migrants_per_substep_rank0': (s.migrated - mig0) / float(steps * nleap),
             'particles_per_rank': [int(c) for c in s.counts],
             'host_ms_per_substep_rank0': {k: 1e3 * t / max(1, (steps + warmup) * nleap) for k, t in s.timing.items()}}
        s.close()
        return ms, d

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
        for _ in range(steps_e):
            s.step(dt_leap, nleap)
            il, xl, vl = s.read_local()
            got = len(il)
        torch.cuda.synchronize()
        el = max_over_ranks(time.perf_counter() - t0)
        s.close()
        return {'value': float(n) * world * nleap * steps_e / el, 'unit': 'particle-steps/s',
                'h2d_bytes_per_step': 20. * n / steps_e, 'd2h_bytes_per_step': 20. * got,
                'note': 'multi.ShardedSystem from host arrays (sample-sort partition + H2D, amortised over %d steps) '
                        '+ step() + read_local() (D2H of x, v, id of the owned particles) each step' % steps_e}

    sharded = world > 1 and a.mode in ('auto', 'sharded')
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    if sharded:
        ms, d = run_sharded(a.dt_leap, a.steps, a.warmup, a.nleap)
    else:
        ms, d = run(a.dt_leap, a.sort, a.steps, a.warmup, a.nleap)
    clocks = sampler.stop() if sampler else None
    psteps = float(n) * world * a.nleap * a.steps
    value = psteps / (ms * 1e-3)

    # N>1, sharded: also time the other multi-GPU mode of the north star -- independent realisations, one per
    # GPU, no data-path collective (BASELINE.json configs[4]) -- so that both scalings can be read off one
    # line.  No collective inside the leg (a rank that fails must not hang the others): local events, then one
    # max over ranks.
    ensemble = None
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

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6650.)
    # dominant kernel: tile_kernel (one launch per sub-step on the bucket path).  Its average
    # duration is the timed region / sub-steps when nothing else ran (rebuilds == 0).
    launches_tile = d['substeps']
    ms_per_launch = ms / max(1, launches_tile)
    achieved = ALG_BYTES * n / (ms_per_launch * 1e-3) / 1e9
    out = {
        'metric': 'particle-steps/s', 'value': value, 'unit': 'particle-steps/s', 'n_gpus': world,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'sech2 disk + harmonic omega=%g, N=%d per GPU, dt_leap=%g, nleap=%d, sort=%s'
                               % (a.omega, n, a.dt_leap, a.nleap, a.sort),
                   'parallelism': ('one system of %d particles range-partitioned over %d GPUs (sample sort: all-to-all of '
                                   'migrants + all-gather of counts per sub-step)' % (n * world, world)) if sharded
                   else ('independent realisations, one per GPU' if world > 1 else 'single GPU'),
                   'l2': 'state (%.1f GB) is far larger than L2' % (n * 28 / 1e9)},
        'gpu_launches': d['kernel_launches'],
        'path_stats': d,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                     'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback', 'unit': 'GB/s',
                     'frac': achieved / peak,
                     # dram__bytes_read+write per launch from the ncu --set full captures under profiles/r01
                     # (wstep: 40.2 B/particle, persistent CTA kernel: 39.9 B/particle -- 2.027 GB read + 1.963 GB
                     # written at N=1e8, tile_1e8_dt1e-3_summary.txt), scaled to this N
                     'traffic': (40.2 if getattr(run, 'cap', 256) == 256 else 39.9) * n if a.sort == 'gpu' else None,
                     'kernel': ('radix passes + tile_kernel<LOAD_GATHER>' if a.sort != 'gpu' else
                                'wstep_kernel<256,8,EQM> (one warp per bucket)' if getattr(run, 'cap', 256) == 256 else
                                'tile_kernel<2048,512,LOAD_BUCKET,EMIT_SPLITTER,EQM,PERSIST> (persistent CTAs, two per SM; next bucket prefetched by TMA)'),
                     'ms_per_launch': ms_per_launch},
        'clocks': clocks,
    }
    if ensemble is not None:
        out['ensemble_mode'] = ensemble

    if a.variants and world == 1:
        var = {}
        for dtl, srt in ((1e-5, 'gpu'), (1e-3, 'gpu'), (5e-3, 'gpu'), (1e-3, 'gpu-radix')):
            vms, vd = run(dtl, srt, max(1, a.steps // 2), 1, a.nleap if srt == 'gpu' else 2)
            nl = a.nleap if srt == 'gpu' else 2
            var['dt_leap=%g,%s' % (dtl, srt)] = {'value': float(n) * nl * max(1, a.steps // 2) / (vms * 1e-3), 'stats': vd}
        # general (unequal) masses: the exact 128-bit scan path
        rs = numpy.random.RandomState(5)
        mj = m * (1. + 0.1 * (2. * rs.uniform(size=n) - 1.))
        for dtl in (1e-5, 1e-3):
            st = wendy_b200.ApproxState(x, v, mj, omega2=omega2, stream=stream)
            st.step(dtl, a.nleap); st.step(dtl, a.nleap)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                st.step(dtl, a.nleap)
            e1.record(); torch.cuda.synchronize()
            var['general masses, dt_leap=%g' % dtl] = {'value': float(n) * a.nleap * 2 / (e0.elapsed_time(e1) * 1e-3),
                                                        'stats': st.stats()}
            st.close()
        # BASELINE config 5 shape: independent realisations of 1e5 particles + torch ext_force (Gaia spiral)
        S, L = max(1, int(n // 100000 // 2)), 100000
        xs = numpy.arctanh(2. * rs.uniform(size=S * L) - 1.) * 2.
        vs = rs.normal(size=S * L) + 1.0
        ms = numpy.full(S * L, 0.3 / L)
        F = lambda xx, t: -0.7 * torch.tanh(0.5 * xx)  # noqa: E731
        g = wendy_b200.nbody(xs, vs, ms, 0.05, approx=True, nleap=10, ext_force=F, n_segments=S)
        next(g)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(2):
            next(g)
        torch.cuda.synchronize(); el = time.perf_counter() - t0
        g.close()
        var['ensemble %d x 1e5 + torch ext_force, dt_leap=0.005 (generator, incl. D2H)' % S] = {'value': S * L * 10 * 2 / el}
        out['variants'] = var

    # ---- end to end through the public generator API, host buffers --------------------------
    if not a.no_e2e and sharded:
        out['e2e'] = e2e_sharded(a.dt_leap, max(2, min(a.steps, 10)), a.nleap)
    if not a.no_e2e and not sharded:
        steps_e = max(2, min(a.steps, 10))
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

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        os.environ.pop('OMP_NUM_THREADS', None)
        nr = int(min(a.ref_n, n))
        val, per, kind = time_reference(x[:nr] if nr < n else x, v[:nr], m[:nr] * (n / nr), a.omega, a.dt_leap, 3, 1)
        out['cpu_baseline'] = {'value': val, 'unit': 'particle-steps/s', 'cores': cores, 'kind': kind,
                               'sample': 'first %d particles of the workload (masses rescaled), 1 warm-up + 3 timed '
                                         'sub-steps, reference sort=parallel, %.2f s per sub-step' % (nr, per)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

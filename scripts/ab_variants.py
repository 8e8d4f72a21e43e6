#!/usr/bin/env python
"""A/B of alternative builds of libwendy_b200.so on one B200 (kernel variants).

  python scripts/ab_variants.py [--install] [--out gpurun_out/ab_variants.json] base.so variant1.so[:cap:fill] ...

Every library runs in its own process (WENDY_B200_LIB): device-generated sech^2 + omega ICs,
N=1e8, 3 warm-up + 5 timed calls of 10 sub-steps at dt_leap = 1e-3, 1e-4, 1e-5 (CUDA events), plus
a parity fingerprint -- sha256 of x and v of an N=3*2^20 system after 20 sub-steps at dt_leap=1e-3
and 10 at 0.05 (window and far-mover paths), which must equal the first library's: every variant
of the step kernel is bit-exact by construction.
--install copies the fastest library whose fingerprints match over wendy_b200/libwendy_b200.so.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DTS = [float(t) for t in os.environ.get("AB_DTS", "1e-3,1e-4,1e-5").split(",")]


def worker(n, cap=0, fill=0):
    import numpy
    import torch
    import wendy_b200
    from wendy_b200 import ic
    out = {'lib': os.environ.get('WENDY_B200_LIB'), 'values': {}, 'ms_per_launch': {}, 'cap': {}}
    # parity fingerprint
    h = hashlib.sha256()
    np_ = 3 << 20
    x, v, m0 = ic.sech2_disk(np_, seed=7)
    for dt, calls in ((1e-3, 2), (0.05, 1)):
        st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21, cap=cap, fill=fill)
        for _ in range(calls):
            st.step(dt, 10)
        xo, vo = st.read()
        out['cap']['parity dt=%g' % dt] = st.stats()['cap']
        out.setdefault('parity_stats', {})['dt=%g' % dt] = st.stats()
        st.close()
        h.update(numpy.ascontiguousarray(xo).tobytes())
        h.update(numpy.ascontiguousarray(vo).tobytes())
    out['fingerprint'] = h.hexdigest()
    del x, v
    # timing
    x, v, m0 = ic.sech2_disk(n, seed=2)
    for dt in DTS:
        st = wendy_b200.ApproxState.from_device(x, v, m0, omega2=1.21, cap=cap, fill=fill)
        for _ in range(3):
            st.step(dt, 10)
        s0 = st.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            st.step(dt, 10)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s1 = st.stats()
        st.close()
        key = 'dt=%g' % dt
        out['values'][key] = n * 50. / (ms * 1e-3)
        out['ms_per_launch'][key] = ms / 50.
        out['cap'][key] = s1['cap']
        out.setdefault('rebuilds', {})[key] = s1['rebuilds'] - s0['rebuilds']
    print('AB_RESULT ' + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('libs', nargs='*')
    ap.add_argument('--one', action='store_true')
    ap.add_argument('--cap', type=int, default=0)
    ap.add_argument('--fill', type=int, default=0)
    ap.add_argument('--install', action='store_true')
    ap.add_argument('--particles', type=float, default=1e8)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'ab_variants.json'))
    a = ap.parse_args()
    if a.one:
        worker(int(a.particles), a.cap, a.fill)
        return
    results = []
    for lib in a.libs:  # "path" or "path:cap:fill" (explicit bucket geometry)
        spec = lib.split(':')
        path = os.path.abspath(spec[0])
        cap, fill = (spec[1], spec[2]) if len(spec) == 3 else ('0', '0')
        env = dict(os.environ, WENDY_B200_LIB=path)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', '--particles', str(a.particles),
                                '--cap', cap, '--fill', fill],
                               env=env, capture_output=True, text=True, timeout=240)
            line = [l for l in p.stdout.splitlines() if l.startswith('AB_RESULT ')]
            r = json.loads(line[-1][len('AB_RESULT '):]) if line else {'error': (p.stderr or p.stdout)[-600:]}
        except subprocess.TimeoutExpired:
            r = {'error': 'timeout'}
        r['lib'] = lib
        results.append(r)
        print(lib, json.dumps({k: r.get(k) for k in ('values', 'fingerprint', 'error', 'cap', 'rebuilds')}), flush=True)
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(results, open(a.out, 'w'), indent=1)
    base = results[0] if results else {}
    best, best_score = None, 0.
    for r in results:
        if 'error' in r or r.get('fingerprint') != base.get('fingerprint'):
            r['accepted'] = False
            continue
        vals = r['values']
        score = 1.
        for vv in vals.values():
            score *= vv ** (1. / len(vals))
        r['score'], r['accepted'] = score, True
        if score > best_score:
            best, best_score = r, score
    summary = {'results': results, 'winner': best['lib'] if best else None}
    json.dump(summary, open(a.out, 'w'), indent=1)
    print('WINNER', summary['winner'])
    if a.install and best and best is not base:
        shutil.copyfile(os.path.abspath(best['lib'].split(':')[0]), os.path.join(ROOT, 'wendy_b200', 'libwendy_b200.so'))
        print('installed', best['lib'])


if __name__ == '__main__':
    main()

"""The bench.py JSON contract.  CPU: the reference arm is run for real on a small sample (it needs no GPU).
GPU: `python bench.py` is run for real at a reduced size and its line is checked key by key -- the line the
driver parses, not a file committed earlier."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline')


def _run(*args, timeout=600):
    env = dict(os.environ)
    env.pop('OMP_NUM_THREADS', None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                       timeout=timeout, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    from oracle import wendy_oracle as wo
    if not wo.reference_available():
        pytest.skip('oracle/_ref not built')
    d = _run('--impl', 'reference', '--particles', '200000', '--steps', '2', '--warmup', '1')
    assert d['impl'] == 'reference'
    for k in BASE_KEYS:
        assert k in d, k
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['e2e']['value'] == d['value'] == d['cpu_baseline']['value'] > 0
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] >= 1
    assert 'N=200000' in d['config']['workload'] and d['config']['host']['nproc'] >= 1
    assert d['unit'] == 'particle-steps/s' and d['higher_is_better'] is True


@pytest.mark.gpu
def test_our_arm_line_carries_every_contract_key_live():
    n = 2000000
    d = _run('--particles', str(n), '--steps', '3', '--warmup', '3')
    for k in BASE_KEYS + ('roofline', 'clocks', 'gpu_launches', 'parity', 'variants', 'path_stats'):
        assert k in d, k
    assert d['unit'] == 'particle-steps/s' and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    assert r['traffic'] is None or r['traffic'] > 0
    # achieved = algorithmic bytes (40 B per particle-step) / launch duration
    assert abs(r['achieved'] - 40. * n / (r['ms_per_launch'] * 1e-3) / 1e9) < 1e-6 * r['achieved']
    assert 'PERSIST=2' in r['kernel'] and d['path_stats']['cap'] == 2048  # the kernel that actually ran
    assert d['path_stats']['substeps'] == 30 and d['gpu_launches'] == d['path_stats']['kernel_launches'] > 0
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert 0 < e['value'] < d['value']  # end to end includes the PCIe copies: never the device-only number
    c = d['cpu_baseline']
    assert c['kind'] in ('reference', 'port') and c['cores'] >= 1 and c['sample']
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    p = d['parity']
    if c['kind'] == 'reference':  # equal masses: bit-identical to the compiled reference on the same system
        for k in ('after_1_substeps', 'after_4_substeps'):
            assert p[k]['x_bit_identical'] and p[k]['v_bit_identical'], p
    assert set(d['variants']) == {'dt_leap=1e-05,gpu', 'dt_leap=0.005,gpu', 'dt_leap=0.001,gpu-radix', 'unequal masses,dt_leap=0.001'}
    for vv in d['variants'].values():
        assert vv.get('value', 0) > 0, vv

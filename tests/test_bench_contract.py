"""The bench.py JSON contract, checked on the lines committed under profiles/ (they were produced by
`python bench.py` and `python bench.py --impl reference` on a B200 box; no GPU is needed to read them)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline')


def _line(name):
    return json.load(open(os.path.join(ROOT, 'profiles', 'r01', name)))


def test_our_arm_line_carries_every_contract_key():
    d = _line('bench_default_N1e8.json')
    for k in BASE_KEYS + ('roofline', 'clocks', 'gpu_launches'):
        assert k in d, k
    assert d['unit'] == 'particle-steps/s' and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    assert r['traffic'] is None or r['traffic'] > 0
    # achieved = algorithmic bytes (40 B per particle-step) / launch duration
    n = 100000000
    assert abs(r['achieved'] - 40. * n / (r['ms_per_launch'] * 1e-3) / 1e9) < 1e-6 * r['achieved']
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert 0 < e['value'] < d['value']  # end to end includes the PCIe copies: never the device-only number
    c = d['cpu_baseline']
    assert c['kind'] in ('reference', 'port') and c['cores'] >= 1 and c['sample']
    assert d['gpu_launches'] > 0
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    bad = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    assert not bad & set(d['clocks']['reasons'])


def test_reference_arm_line():
    d = _line('bench_reference.json')
    assert d['impl'] == 'reference'
    for k in BASE_KEYS:
        assert k in d, k
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['e2e']['value'] == d['value'] == d['cpu_baseline']['value']
    assert d['cpu_baseline']['kind'] == 'reference'

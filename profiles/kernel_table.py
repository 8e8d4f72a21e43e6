"""ncu --csv launch list (gpu__time_duration.sum, dram bytes, issue/warps active) -> per-kernel table.
   python profiles/kernel_table.py launches.csv [peak_GBs] > profiles/r02/kernels.md"""
import csv
import re
import sys
from collections import OrderedDict

peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6558.1
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
I = {k: hdr.index(k) for k in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value')}
launch = OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    d = launch.setdefault(r[I['ID']], {'name': r[I['Kernel Name']]})
    val = float(r[I['Metric Value']].replace(',', '') or 0)
    unit = r[I['Metric Unit']]
    if unit in ('ns', 'nsecond'):
        val /= 1e6
    elif unit in ('us', 'usecond'):
        val /= 1e3
    elif unit in ('s', 'second'):
        val *= 1e3
    elif unit == 'Kbyte':
        val *= 1e3
    elif unit == 'Mbyte':
        val *= 1e6
    elif unit == 'Gbyte':
        val *= 1e9
    d[r[I['Metric Name']]] = val


def short(n):
    n = re.sub(r'^void\s+', '', n)
    n = n.replace('wendy::', '')
    n = re.sub(r'\((?:[^()]|\([^()]*\))*\)$', '', n)
    return n[:90]

agg = OrderedDict()
for d in launch.values():
    a = agg.setdefault(short(d['name']), {'n': 0, 'ms': 0., 'max_ms': 0., 'bytes_at_max': 0., 'issue': 0., 'warps': 0., 'regs': 0})
    ms = d.get('gpu__time_duration.sum', 0.)
    a['n'] += 1
    a['ms'] += ms
    if ms >= a['max_ms']:
        a['max_ms'] = ms
        a['bytes_at_max'] = d.get('dram__bytes_read.sum', 0.) + d.get('dram__bytes_write.sum', 0.)
        a['issue'] = d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0.)
        a['warps'] = d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0.)
        a['regs'] = d.get('launch__registers_per_thread', 0)
print('| kernel | launches | total ms | longest launch ms | DRAM bytes of that launch | GB/s | %% of %.0f GB/s | issue active %% | warps active %% | regs |' % peak)
print('|---|---|---|---|---|---|---|---|---|---|')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
    gbs = a['bytes_at_max'] / (a['max_ms'] * 1e-3) / 1e9 if a['max_ms'] > 0 else 0.
    print('| `%s` | %d | %.3f | %.4f | %.3e | %.0f | %.1f | %.0f | %.0f | %d |' % (k, a['n'], a['ms'], a['max_ms'], a['bytes_at_max'], gbs, 100 * gbs / peak, a['issue'], a['warps'], a['regs']))

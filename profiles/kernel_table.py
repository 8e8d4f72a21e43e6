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
    a = agg.setdefault(short(d['name']), [])
    a.append(d)
print('| kernel | launches | total ms | representative launch ms | DRAM bytes of that launch | GB/s | %% of %.0f GB/s | issue active %% | warps active %% | regs |' % peak)
print('|---|---|---|---|---|---|---|---|---|---|')
rows_out = []
for k, ds in agg.items():
    # representative launch: among the launches at least half as long as the longest (the full-size ones; the tour
    # also runs smaller systems), the fastest per byte -- the first launch of a kind runs on cold caches under ncu
    tmax = max(d.get('gpu__time_duration.sum', 0.) for d in ds)
    big = [d for d in ds if d.get('gpu__time_duration.sum', 0.) >= 0.5 * tmax]
    def rate(d):
        t = d.get('gpu__time_duration.sum', 0.)
        return (d.get('dram__bytes_read.sum', 0.) + d.get('dram__bytes_write.sum', 0.)) / t if t > 0 else 0.
    m = max(big, key=rate)
    ms = m.get('gpu__time_duration.sum', 0.)
    by = m.get('dram__bytes_read.sum', 0.) + m.get('dram__bytes_write.sum', 0.)
    gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.
    tot = sum(d.get('gpu__time_duration.sum', 0.) for d in ds)
    rows_out.append((tot, '| `%s` | %d | %.3f | %.4f | %.3e | %.0f | %.1f | %.0f | %.0f | %d |' % (
        k, len(ds), tot, ms, by, gbs, 100 * gbs / peak, m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0.),
        m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0.), m.get('launch__registers_per_thread', 0))))
for _, line in sorted(rows_out, key=lambda t: -t[0]):
    print(line)

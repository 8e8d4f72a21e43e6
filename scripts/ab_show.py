#!/usr/bin/env python
"""Print the table of an ab_variants.json (scripts/ab_variants.py)."""
import json
import sys
d = json.load(open(sys.argv[1]))
rs = d['results'] if isinstance(d, dict) else d
b = rs[0]
for r in rs:
    if r.get('values'):
        print('%-44s %s fp_ok=%s score=%.4g' % (r['lib'].replace('wendy_b200/variants/', ''), ' '.join(
            '%s %.3e (%+5.1f%%)' % (k, v, 100 * (v / b['values'][k] - 1)) for k, v in r['values'].items()),
            r['fingerprint'] == b['fingerprint'], r.get('score', 0)))
    else:
        print(r['lib'], (r.get('error') or '')[-300:])

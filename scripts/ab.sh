#!/bin/bash
# A/B of kernel variants on one box: scripts/ab.sh <lib1.so> <lib2.so> ...   (paths relative to the repo root)
# Prints particle-steps/s of the default bench (device-resident, N=1e8) at dt_leap = 1e-3 and 1e-5 per library.
for lib in "$@"; do
  for dt in 1e-3 1e-5; do
    WENDY_B200_LIB=$PWD/$lib timeout 300 python bench.py --skip-cpu-baseline --skip-e2e --dt-leap $dt --steps 10 --warmup 3 2>/dev/null \
      | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib dt=$dt value %.4e  ms/launch %.4f  %s' % (d['value'], d['roofline']['ms_per_launch'], d['roofline']['kernel'][:24]))"
  done
done

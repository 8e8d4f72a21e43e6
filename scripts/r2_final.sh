#!/bin/bash
# round 2, last call (1 GPU): GPU suite, default bench line, full ncu capture of the step kernel -- each under its own timeout
O=gpurun_out/r2fin
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 110 python -m pytest tests -x -q -m gpu -n 4 > $O/pytest_gpu.log 2>&1
echo "pytest: $(tail -1 $O/pytest_gpu.log) $(el)"
grep -E "^FAILED|^ERROR" $O/pytest_gpu.log | head -5
timeout 110 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(el)"; python -c "import json; d=json.load(open('$O/bench_default.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], {k: v.get('value') for k, v in d.get('variants', {}).items()}, d.get('parity', {}).get('after_4_substeps'))"
timeout 70 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 25 -c 1 \
    -o $O/tile_1e8_dt1e-3 -f python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-variants > $O/p1.log 2>&1
echo "ncu full exit $? $(el)"

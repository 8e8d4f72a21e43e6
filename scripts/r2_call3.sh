#!/bin/bash
# round 2, call 3: onesweep check, GPU suite (names, per-test timeout), multi-GPU exchange, micro-benchmark, A/B, bench, launch list
O=gpurun_out/r2c3
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 120 python -m pytest tests/test_gpu_parity.py -k "argsort" -q -x > $O/pytest_argsort.log 2>&1
echo "pytest(argsort, onesweep) exit $? $(tail -1 $O/pytest_argsort.log) $(el)"
if ! tail -1 $O/pytest_argsort.log | grep -q " passed"; then
  echo "ONESWEEP BROKEN -> WENDY_B200_RADIX=lsd for the rest"; tail -30 $O/pytest_argsort.log
  export WENDY_B200_RADIX=lsd
fi
timeout 700 python -m pytest tests -m gpu -v -n 4 --timeout 150 --durations=12 --ignore tests/test_gpu_reference_scale.py --ignore tests/test_gpu_multi.py > $O/pytest_gpu.log 2>&1
echo "pytest(main) exit $? $(tail -1 $O/pytest_gpu.log) $(el)"
grep -E "FAILED|ERROR|Timeout" $O/pytest_gpu.log | sort | uniq | head -30
timeout 400 python -m pytest tests/test_gpu_multi.py -v --timeout 120 --durations=10 > $O/pytest_multi.log 2>&1
echo "pytest(multi) exit $? $(tail -1 $O/pytest_multi.log) $(el)"
grep -E "PASSED|FAILED|ERROR|Error|SKIPPED" $O/pytest_multi.log | cut -c1-200 | head -40
timeout 60 scripts/ubench/atoms > $O/atoms.txt 2>&1; cat $O/atoms.txt
V=wendy_b200/variants
timeout 420 python scripts/ab_variants.py --out $O/ab_variants.json $V/lib_base.so $V/lib_exact.so $V/lib_lazy.so $V/lib_sm2.so $V/lib_sm2l.so $V/lib_sm2lr2.so $V/lib_lazyr2.so > $O/ab.log 2>&1
cut -c1-260 $O/ab.log
echo "ab done $(el)"
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(el)"; tail -3 $O/bench_default.err
head -c 6000 $O/bench_default.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 2 --skip-e2e --skip-cpu-baseline > $O/launches_default.log 2>&1
echo "launch list exit $? $(el)"

#!/bin/bash
# round 2, call 5 (1 GPU): the driver's GPU test command, emission/ranking A/B, default bench, per-kernel ncu tour
O=gpurun_out/r2c5
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 1200 python -m pytest tests -x -q -m gpu --durations=15 > $O/pytest_gpu.log 2>&1
echo "pytest(driver command) exit $? $(tail -1 $O/pytest_gpu.log) $(el)"
grep -E "FAILED|ERROR|Timeout" $O/pytest_gpu.log | sort | uniq | head -20
V=wendy_b200/variants
timeout 500 python scripts/ab_variants.py --out $O/ab_variants.json $V/lib_base.so $V/lib_old.so $V/lib_em1.so $V/lib_em2.so $V/lib_r3.so $V/lib_r1.so > $O/ab.log 2>&1
cut -c1-200 $O/ab.log
echo "ab done $(el)"
timeout 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(el)"; tail -3 $O/bench_default.err
head -c 7000 $O/bench_default.json
timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none --csv --log-file $O/tour.csv python scripts/kernel_tour.py > $O/tour.log 2>&1
echo "tour exit $? $(el)"; tail -3 $O/tour.log

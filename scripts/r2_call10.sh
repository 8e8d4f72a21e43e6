#!/bin/bash
# round 2, call 10 (1 GPU): final single-GPU records -- driver's test command, default bench, config legs, reference arm, launch list, per-kernel tour, full capture of the step kernel
O=gpurun_out/r2c10
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 > $O/pytest_gpu.log 2>&1
echo "pytest(driver command) exit $? $(tail -1 $O/pytest_gpu.log) $(el)"
grep -E "FAILED|ERROR|Timeout" $O/pytest_gpu.log | sort | uniq | head -20
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $? $(tail -1 $O/smoke.log) $(el)"
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(el)"; tail -3 $O/bench_default.err; head -c 1500 $O/bench_default.json; echo
for c in 1 2 5; do
  timeout 600 python bench.py --config $c --steps 10 > $O/bench_config$c.json 2> $O/bench_config$c.err
  echo "config $c exit $? $(el)"; tail -2 $O/bench_config$c.err | cut -c1-300; head -c 700 $O/bench_config$c.json; echo
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "reference arm exit $? $(el)"; head -c 1800 $O/bench_reference.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline > $O/launches_default.log 2>&1
echo "launch list exit $? $(el)"
timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none --csv --log-file $O/tour.csv python scripts/kernel_tour.py > $O/tour.log 2>&1
echo "tour exit $? $(el)"; tail -2 $O/tour.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 25 -c 1 \
    -o $O/tile_1e8_dt1e-3 -f python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-variants > $O/p1.log 2>&1
echo "ncu full exit $? $(el)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 3 -c 1 \
    -o $O/onesweep_1e8 -f python bench.py --steps 1 --warmup 1 --skip-e2e --skip-cpu-baseline --skip-variants > $O/p2.log 2>&1
echo "ncu onesweep exit $? $(el)"

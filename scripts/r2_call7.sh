#!/bin/bash
# round 2, call 7 (1 GPU): driver's GPU test command, shard-instance diagnosis, BASELINE configs 1/2/5 legs, default bench, full ncu capture of the step kernel
O=gpurun_out/r2c7
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > $O/pytest_gpu.log 2>&1
echo "pytest(driver command) exit $? $(tail -1 $O/pytest_gpu.log) $(el)"
grep -E "FAILED|ERROR|Timeout" $O/pytest_gpu.log | sort | uniq | head -20
timeout 300 python scripts/shard_solo.py > $O/shard_solo.log 2>&1; echo "solo exit $? $(el)"; grep -E "instance|Error|error" $O/shard_solo.log | cut -c1-300
for c in 1 2 5; do
  timeout 600 python bench.py --config $c --steps 5 > $O/bench_config$c.json 2> $O/bench_config$c.err
  echo "config $c exit $? $(el)"; tail -2 $O/bench_config$c.err | cut -c1-300; head -c 2500 $O/bench_config$c.json; echo
done
timeout 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(el)"; tail -3 $O/bench_default.err
head -c 7000 $O/bench_default.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 25 -c 1 \
    -o $O/tile_1e8_dt1e-3 -f python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-variants > $O/p1.log 2>&1
echo "ncu full exit $? $(el)"

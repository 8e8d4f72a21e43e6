#!/bin/bash
# One gpurun call (1 GPU) that does, in priority order and each under its own timeout:
#   1. the GPU parity suite
#   2. the default bench line (and the reference arm)
#   3. the ncu launch list of the default bench command
#   4. one ncu --set full capture of the dominant kernel
#   5. compute-sanitizer racecheck/memcheck on a small case
#   6. phase timings of the end-to-end generator path
#   7. A/B of the variant builds wendy_b200/variants/lib2_*.so against the default build (scripts/ab_variants.py)
# Everything lands under gpurun_out/round/.
O=gpurun_out/round
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
date +%s > $O/t0
timeout 420 python -m pytest tests -m gpu -x -q -n 4 > $O/pytest_gpu.log 2>&1
echo "pytest(xdist) exit $? $(tail -1 $O/pytest_gpu.log) $(( $(date +%s) - $(cat $O/t0) )) s"
if ! tail -1 $O/pytest_gpu.log | grep -q " passed"; then
  timeout 420 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_serial.log 2>&1
  echo "pytest(serial) exit $? $(tail -1 $O/pytest_gpu_serial.log)"
fi
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(( $(date +%s) - $(cat $O/t0) )) s"
head -c 1500 $O/bench_default.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline > $O/launches_default.log 2>&1
echo "launch list exit $? $(( $(date +%s) - $(cat $O/t0) )) s"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 25 -c 1 \
    -o $O/tile_1e8_dt1e-3 -f python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline > $O/p1.log 2>&1
echo "ncu full exit $? $(( $(date +%s) - $(cat $O/t0) )) s"
timeout 200 python bench.py --dt-leap 1e-5 --skip-e2e --skip-cpu-baseline > $O/bench_dt1e-5.json 2> $O/bench_dt1e-5.err
timeout 240 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $O/racecheck.log 2>&1
echo "racecheck exit $? $(tail -2 $O/racecheck.log | head -1)"
timeout 240 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $O/memcheck.log 2>&1
echo "memcheck exit $? $(tail -2 $O/memcheck.log | head -1)"
timeout 120 python scripts/e2e_phases.py > $O/e2e_phases.log 2>&1
cat $O/e2e_phases.log
WENDY_B200_TRACE=1 timeout 120 python scripts/e2e_phases.py > $O/e2e_phases_trace.log 2>&1
WENDY_B200_PIN_CHUNK_MB=65536 WENDY_B200_TRACE=1 timeout 120 python scripts/e2e_phases.py > $O/e2e_phases_trace_unchunked.log 2>&1
grep "generator" $O/e2e_phases_trace.log $O/e2e_phases_trace_unchunked.log
# A/B of whatever variant builds are present (never installed here: the stages above test the default build)
cp wendy_b200/libwendy_b200.so wendy_b200/variants/lib_base.so
VARS="wendy_b200/variants/lib_base.so"
for f in wendy_b200/variants/lib2_*.so; do [ -f $f ] && VARS="$VARS $f"; done
timeout 300 python scripts/ab_variants.py --out $O/ab_variants.json $VARS > $O/ab.log 2>&1
tail -3 $O/ab.log
ls -la $O | tail -20

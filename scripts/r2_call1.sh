#!/bin/bash
# round 2, call 1: GPU suite (incl. parity vs the compiled reference at 1e7/1e8), A/B of the prepared kernel variants, bench
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; free -g >> $O/nproc.txt
date +%s > $O/t0
timeout 700 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_gpu_reference_scale.py > $O/pytest_gpu.log 2>&1
echo "pytest(xdist) exit $? $(tail -1 $O/pytest_gpu.log) $(( $(date +%s) - $(cat $O/t0) )) s"
grep -E "FAILED|ERROR" $O/pytest_gpu.log | head -20
timeout 600 python -m pytest tests/test_gpu_reference_scale.py -q -x --durations=5 > $O/pytest_scale.log 2>&1
echo "pytest(scale) exit $? $(tail -1 $O/pytest_scale.log) $(( $(date +%s) - $(cat $O/t0) )) s"
grep -E "FAILED|ERROR|assert" $O/pytest_scale.log | head -20
V=wendy_b200/variants
timeout 600 python scripts/ab_variants.py --out $O/ab_variants.json $V/lib_base.so $V/lib_exact.so $V/lib_os.so $V/lib_rd.so $V/lib_rdos.so $V/lib_pe8.so $V/lib_os8.so $V/lib_wp.so $V/lib_nw.so $V/lib_s32.so $V/lib_all3.so > $O/ab.log 2>&1
cat $O/ab.log | cut -c1-400
echo "ab done $(( $(date +%s) - $(cat $O/t0) )) s"
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench exit $? $(( $(date +%s) - $(cat $O/t0) )) s"
head -c 2500 $O/bench_default.json

#!/bin/bash
# round 2, call 2: GPU suite by parts (names + durations), multi-GPU exchange as threads on one GPU
O=gpurun_out/r2c2
mkdir -p $O
date +%s > $O/t0
timeout 600 python -m pytest tests -m gpu -q -n 4 --durations=15 --ignore tests/test_gpu_reference_scale.py --ignore tests/test_gpu_multi.py > $O/pytest_gpu.log 2>&1
echo "pytest(main) exit $? $(tail -1 $O/pytest_gpu.log) $(( $(date +%s) - $(cat $O/t0) )) s"
grep -E "^FAILED|^ERROR" $O/pytest_gpu.log | head -20
timeout 400 python -m pytest tests/test_gpu_multi.py -v -x --durations=10 > $O/pytest_multi.log 2>&1
echo "pytest(multi) exit $? $(tail -1 $O/pytest_multi.log) $(( $(date +%s) - $(cat $O/t0) )) s"
grep -E "PASSED|FAILED|ERROR|Error" $O/pytest_multi.log | head -30
timeout 600 python -m pytest tests/test_gpu_reference_scale.py -v --durations=5 > $O/pytest_scale.log 2>&1
echo "pytest(scale) exit $? $(tail -1 $O/pytest_scale.log) $(( $(date +%s) - $(cat $O/t0) )) s"
grep -E "PASSED|FAILED|ERROR|assert" $O/pytest_scale.log | head -20

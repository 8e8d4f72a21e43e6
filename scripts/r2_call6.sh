#!/bin/bash
# round 2, call 6 (2 GPUs): multi tests, sharded bench with the per-kernel trace
O=gpurun_out/r2c6
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q --timeout 120 --durations=5 > $O/pytest_multi.log 2>&1
echo "pytest(multi) exit $? $(tail -1 $O/pytest_multi.log) $(el)"
grep -E "FAILED|ERROR|Error" $O/pytest_multi.log | cut -c1-200 | head -10
WENDY_B200_SHARD_TRACE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2.json 2> $O/bench_2.err
echo "bench(2) exit $? $(el)"; grep -v "^\*\*\*\|OMP_NUM" $O/bench_2.err | tail -12 | cut -c1-400
grep '^{' $O/bench_2.json | head -c 5000

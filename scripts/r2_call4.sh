#!/bin/bash
# round 2, call 4 (2 GPUs): the device-driven exchange between two PROCESSES (CUDA IPC over NVLink), NCCL test, bench at N=2
O=gpurun_out/r2c4
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -v --timeout 120 --durations=10 -k "nccl or four_ranks or rolls_back" > $O/pytest_multi.log 2>&1
echo "pytest(multi) exit $? $(tail -1 $O/pytest_multi.log) $(el)"
grep -E "PASSED|FAILED|ERROR|Error|SKIPPED" $O/pytest_multi.log | cut -c1-200 | head -40
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/shard_probe.py 2e6 > $O/probe.log 2>&1
echo "probe exit $? $(el)"; tail -12 $O/probe.log | cut -c1-300
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2.json 2> $O/bench_2.err
echo "bench(2) exit $? $(el)"; tail -5 $O/bench_2.err | cut -c1-300
grep '^{' $O/bench_2.json | head -c 5000

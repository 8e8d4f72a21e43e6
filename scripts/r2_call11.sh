#!/bin/bash
# round 2, final multi-GPU records: bench.py --gpus N (sharded system + regime variant + ensemble leg + sharded_check + e2e); N=8 also config 5 at shape and N=1e9
G=${1:-2}
O=gpurun_out/r2c11_$G
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
WENDY_B200_SHARD_TRACE=1 timeout 600 $TR --master-port 29531 bench.py --gpus $G --steps 10 --warmup 3 > $O/bench_$G.json 2> $O/bench_$G.err
echo "bench($G) exit $? $(el)"; grep -E "shard trace" $O/bench_$G.err | grep " 100 sub-steps" | cut -c1-200 | head -8
grep '^{' $O/bench_$G.json | head -c 4500; echo
if [ "$G" = "2" ]; then
  timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k nccl > $O/pytest_nccl.log 2>&1; echo "nccl test: $(tail -1 $O/pytest_nccl.log) $(el)"
fi
if [ "$G" = "8" ]; then
  timeout 400 $TR --master-port 29532 bench.py --gpus $G --config 5 --steps 10 > $O/bench_config5_$G.json 2> $O/bench_config5_$G.err
  echo "config5($G) exit $? $(el)"; grep '^{' $O/bench_config5_$G.json | head -c 1200; echo
  timeout 500 $TR --master-port 29533 bench.py --gpus $G --particles 1.25e8 --steps 5 --warmup 2 --skip-e2e > $O/bench_1e9_$G.json 2> $O/bench_1e9_$G.err
  echo "bench N=1e9($G) exit $? $(el)"; grep '^{' $O/bench_1e9_$G.json | head -c 3000; echo
fi

#!/bin/bash
# round 2, call 9 (8 GPUs): the driver's scaling command at N=8, config 4 as written (N=1e9), config 5 at shape (4096 x 1e5)
O=gpurun_out/r2c9
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
G=${1:-8}
nproc > $O/nproc.txt; free -g | head -2 >> $O/nproc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
WENDY_B200_SHARD_TRACE=1 timeout 500 $TR --master-port 29521 bench.py --gpus $G --steps 10 --warmup 3 > $O/bench_$G.json 2> $O/bench_$G.err
echo "bench($G) exit $? $(el)"; grep -E "shard trace|Error|error" $O/bench_$G.err | cut -c1-330 | head -20
grep '^{' $O/bench_$G.json | head -c 6000; echo
timeout 400 $TR --master-port 29522 bench.py --gpus $G --config 5 --steps 5 > $O/bench_config5_$G.json 2> $O/bench_config5_$G.err
echo "config5($G) exit $? $(el)"; grep -E "Error|error" $O/bench_config5_$G.err | head -5; grep '^{' $O/bench_config5_$G.json | head -c 1500; echo
timeout 500 $TR --master-port 29523 bench.py --gpus $G --particles 1.25e8 --steps 5 --warmup 2 --skip-e2e > $O/bench_1e9_$G.json 2> $O/bench_1e9_$G.err
echo "bench N=1e9($G) exit $? $(el)"; grep -E "Error|error" $O/bench_1e9_$G.err | head -5; grep '^{' $O/bench_1e9_$G.json | head -c 5000; echo

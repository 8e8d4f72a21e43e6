#!/bin/bash
# round 2, call 8 (1 GPU): shard instance after the edge-handling move, config 5 probe, config 2 with the radix variant, default bench (kernel only)
O=gpurun_out/r2c8
mkdir -p $O
date +%s > $O/t0
el() { echo "$(( $(date +%s) - $(cat $O/t0) )) s"; }
timeout 300 python scripts/shard_solo.py > $O/shard_solo.log 2>&1; echo "solo exit $? $(el)"; grep -E "instance|Error|error" $O/shard_solo.log | cut -c1-300
timeout 300 python scripts/config5_probe.py 128 6 > $O/config5_probe.log 2>&1; echo "probe exit $? $(el)"; tail -14 $O/config5_probe.log | cut -c1-300
timeout 300 python bench.py --config 2 --skip-cpu-baseline > $O/bench_config2.json 2> $O/bench_config2.err
echo "config 2 exit $? $(el)"; tail -2 $O/bench_config2.err | cut -c1-300; head -c 1500 $O/bench_config2.json; echo
timeout 300 python bench.py --skip-cpu-baseline --skip-e2e > $O/bench_kernel.json 2> $O/bench_kernel.err
echo "bench exit $? $(el)"; tail -3 $O/bench_kernel.err; head -c 2500 $O/bench_kernel.json
timeout 200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $? $(tail -1 $O/pytest.log) $(el)"

#!/bin/bash
# Short gpurun call: GPU suite and the default bench line of the tree as committed
O=gpurun_out/final2
mkdir -p $O
timeout 50 python -m pytest tests -m gpu -x -q -n 4 > $O/pytest_gpu.log 2>&1; echo "pytest: $(tail -1 $O/pytest_gpu.log)"
timeout 50 python bench.py > $O/bench_default.json 2> $O/bench_default.err; python -c "import json; d=json.load(open('$O/bench_default.json')); print('bench', d['value'], d['e2e']['value'])"

#!/bin/bash
# Short gpurun call: GPU suite and default bench in both read-out modes (bounce-buffered, the default, and
# WENDY_B200_D2H=pinned: page-locked yield buffers)
O=gpurun_out/final
mkdir -p $O
timeout 60 python -m pytest tests -m gpu -x -q -n 4 > $O/pytest_bounce.log 2>&1; echo "pytest bounce: $(tail -1 $O/pytest_bounce.log)"
WENDY_B200_D2H=pinned timeout 60 python -m pytest tests -m gpu -x -q -n 4 > $O/pytest_pinned.log 2>&1; echo "pytest pinned: $(tail -1 $O/pytest_pinned.log)"
timeout 60 python bench.py > $O/bench_bounce.json 2> $O/bench_bounce.err; python -c "import json; d=json.load(open('$O/bench_bounce.json')); print('bounce', d['value'], d['e2e']['value'])"
WENDY_B200_D2H=pinned timeout 60 python bench.py --skip-cpu-baseline > $O/bench_pinned.json 2> $O/bench_pinned.err; python -c "import json; d=json.load(open('$O/bench_pinned.json')); print('pinned', d['value'], d['e2e']['value'])"
WENDY_B200_TRACE=1 timeout 40 python scripts/e2e_phases.py > $O/e2e_phases_trace.log 2>&1; grep "generator\|read_end" $O/e2e_phases_trace.log | head -4

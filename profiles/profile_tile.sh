#!/bin/bash
# Run under gpurun (1 GPU).  Produces the launch list and one full capture of the hot kernel.
set -x
N=${N:-2e7}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --particles $N --steps 2 --warmup 1 --skip-e2e --skip-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 12 -c 2 \
    -o gpurun_out/tile_prof -f python bench.py --particles $N --steps 2 --warmup 1 --skip-e2e --skip-cpu-baseline > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out

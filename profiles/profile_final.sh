#!/bin/bash
# Run under gpurun (1 GPU).  Launch list of the default bench command + full captures of the step kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline > gpurun_out/launches_default.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 25 -c 1 \
    -o gpurun_out/tile_1e8_dt1e-3 -f python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 25 -c 1 \
    -o gpurun_out/tile_1e8_dt1e-5 -f python bench.py --dt-leap 1e-5 --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wstep_kernel -s 25 -c 1 \
    -o gpurun_out/wstep_1e8_dt1e-5 -f python bench.py --dt-leap 1e-5 --cap 256 --steps 1 --warmup 3 --skip-e2e --skip-cpu-baseline > gpurun_out/p3.log 2>&1
ls -la gpurun_out | tail -8

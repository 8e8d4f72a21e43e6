#!/bin/bash
# Run under gpurun (1 GPU): full ncu capture of the warp-per-bucket kernel at two time steps.
N=${N:-2e7}
mkdir -p gpurun_out
for DT in 1e-5 1e-3; do
ncu --set full --clock-control none --import-source on -k regex:wstep_kernel -s 12 -c 1 \
    -o gpurun_out/wstep_$DT -f python bench.py --particles $N --dt-leap $DT --steps 2 --warmup 1 --skip-e2e --skip-cpu-baseline > gpurun_out/prof_bench_$DT.log 2>&1
done
ls -la gpurun_out

#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench run + one full-set capture of the top kernel.
# Numbers printed by a run under ncu are never bench values.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_locate_points -s 3 -c 1 -f -o gpurun_out/locate_points_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out

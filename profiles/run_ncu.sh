#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench run + full-set captures of the two top kernels
# (point traversal on C2, cooperative segment walk on C4).  Numbers printed by a run under ncu are never bench values.
# Read the reports offline with `python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep`.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_locate_points -s 3 -c 1 -f -o gpurun_out/locate_points_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
NQ=10000000 ncu --set full --clock-control none --import-source on -k regex:k_edges_cooperative -s 1 -c 1 -f \
    -o gpurun_out/edges_cooperative_${TAG} python profiles/exp_edges.py > gpurun_out/edges_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out

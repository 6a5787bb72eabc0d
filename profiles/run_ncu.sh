#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench run + full-set captures of the kernels of the point path
# (binning count / scatter, tile traversal, windows) on C2 and of the cooperative segment walk on C4.
# Numbers printed by a run under ncu are never bench values.
# Read the reports offline with `python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep`.
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
WEIGHTS=0 ncu --set full --clock-control none --import-source on -k regex:"k_locate_points_binned|k_slab_scatter|k_locate_points_overflow|k_windows_to_out" -s 8 -c 4 -f \
    -o gpurun_out/points_${TAG} python profiles/exp_points.py > gpurun_out/points_under_ncu_${TAG}.log 2>&1
NQ=10000000 ncu --set full --clock-control none --import-source on -k regex:k_edges_cooperative -s 1 -c 1 -f \
    -o gpurun_out/edges_cooperative_${TAG} python profiles/exp_edges.py > gpurun_out/edges_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out

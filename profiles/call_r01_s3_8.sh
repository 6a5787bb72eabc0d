set -x
mkdir -p gpurun_out
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -n 6 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python profiles/exp_c3.py > gpurun_out/exp_c3.log 2>&1; cat gpurun_out/exp_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_locate_points -s 3 -c 1 -f -o gpurun_out/locate_points_s3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu_full.log 2>&1
NQ=10000000 ncu --set full --clock-control none --import-source on -k regex:k_edges_cooperative -s 1 -c 1 -f -o gpurun_out/edges_cooperative_s3 python profiles/exp_edges.py > gpurun_out/edges_ncu_full.log 2>&1
ls -la gpurun_out

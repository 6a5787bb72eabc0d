set -x
mkdir -p gpurun_out
NQ=10000000 python profiles/exp_edges.py > gpurun_out/edges_alone.log 2>&1; cat gpurun_out/edges_alone.log
NQ=10000000 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_edges.csv python profiles/exp_edges.py > gpurun_out/edges_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_edges.csv | head -12
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -n 4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.err

set -x
mkdir -p gpurun_out
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -n 6 gpurun_out/pytest_gpu.log
NQ=10000000 python profiles/exp_edges.py > gpurun_out/edges_rank.log 2>&1; cat gpurun_out/edges_rank.log
NQ=10000000 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_edges_rank.csv python profiles/exp_edges.py > gpurun_out/edges_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_edges_rank.csv > gpurun_out/launches_edges_rank.txt; head -n 6 gpurun_out/launches_edges_rank.txt

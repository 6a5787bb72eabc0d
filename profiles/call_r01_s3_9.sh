set -x
mkdir -p gpurun_out
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -n 6 gpurun_out/pytest_gpu.log
NQ=10000000 python profiles/exp_edges.py > gpurun_out/edges_nodiv.log 2>&1; cat gpurun_out/edges_nodiv.log
python profiles/exp_e2e.py > gpurun_out/e2e_check.log 2>&1; tail -n 2 gpurun_out/e2e_check.log

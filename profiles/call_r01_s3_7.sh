set -x
mkdir -p gpurun_out
( time timeout 540 python -m pytest tests -m gpu -x -q -k "edge or golden or kat or smoke" ) > gpurun_out/pytest_gpu_edges.log 2>&1; tail -n 6 gpurun_out/pytest_gpu_edges.log
for r in 1 2 4 8; do CELLTREE_EDGE_ROUNDS=$r NQ=10000000 python profiles/exp_edges.py > gpurun_out/edges_rounds_$r.log 2>&1; tail -n 1 gpurun_out/edges_rounds_$r.log; done

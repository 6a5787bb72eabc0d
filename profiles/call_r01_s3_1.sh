set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 600 gpurun_out/bench_1gpu.err
CELLTREE_HOST_TAIL=0 python profiles/exp_e2e.py > gpurun_out/e2e_tail0.log 2>&1
python profiles/exp_e2e.py > gpurun_out/e2e_tail_default.log 2>&1
CELLTREE_HOST_TAIL=65536 python profiles/exp_e2e.py > gpurun_out/e2e_tail_64k.log 2>&1
tail -2 gpurun_out/e2e_*.log

set -x
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; tail -n 15 gpurun_out/pytest_multi.log
for b in 16 24; do CELLTREE_SORT_BITS=$b WEIGHTS=0 python profiles/exp_points.py > gpurun_out/sortbits_c2_$b.log 2>&1; tail -n 1 gpurun_out/sortbits_c2_$b.log; done

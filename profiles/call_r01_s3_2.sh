set -x
mkdir -p gpurun_out
for b in 10 11 12; do CELLTREE_ENTRY_BITS=$b WEIGHTS=0 python profiles/exp_points.py > gpurun_out/entry_c2_$b.log 2>&1; tail -n 1 gpurun_out/entry_c2_$b.log; done
for b in 7 9 10 11; do CELLTREE_ENTRY_BITS=$b python profiles/exp_c3_points.py > gpurun_out/entry_c3_$b.log 2>&1; tail -n 1 gpurun_out/entry_c3_$b.log; done
python profiles/exp_c3.py > gpurun_out/exp_c3.log 2>&1; cat gpurun_out/exp_c3.log

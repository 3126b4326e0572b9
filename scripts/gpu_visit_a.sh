#!/bin/bash
# visit A: parity tests + smoke + bench + steady-state ncu capture (c_bar ~ 9) of the force kernel with source page
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -5 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_force --launch-skip 1300 -c 1 \
    -f -o $OUT/force python bench.py --steps 1 --warmup 3 --substeps 350 --cpu-steps 0 > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT

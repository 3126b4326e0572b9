#!/bin/bash
TAG=${1:-n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_force -s 40 -c 1 \
    -f -o $OUT/force python bench.py --steps 1 --warmup 3 --substeps 20 --cpu-steps 0 > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT

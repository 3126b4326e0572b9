#!/bin/bash
# ncu visit: launch list + full capture of the force kernel.  usage: bash scripts/gpu_ncu.sh <tag> [extra bench args]
TAG=${1:-n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 270 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --substeps 20 --cpu-steps 0 ${@:2} > $OUT/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_force -s 40 -c 2 \
    -f -o $OUT/force python bench.py --steps 1 --warmup 3 --substeps 20 --cpu-steps 0 ${@:2} > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT

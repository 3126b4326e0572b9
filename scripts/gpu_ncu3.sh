#!/bin/bash
# steady-state capture of the force kernel (launch ~1300 of the bench run: c_bar ~9) + launch list of a steady step
TAG=${1:-n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_force -s 1300 -c 1 \
    -f -o $OUT/force python bench.py --steps 11 --warmup 3 --substeps 100 --cpu-steps 0 > $OUT/ncu_full_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11700 -c 180 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 11 --warmup 3 --substeps 100 --cpu-steps 0 > $OUT/ncu_launch_bench.log 2>&1
ls -la $OUT

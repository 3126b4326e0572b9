#!/bin/bash
# full capture of one REAL rebuild launch of k_build_list (skin 0: every step rebuilds)
OUT=gpurun_out/${1:-nl}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_build_list -s 60 -c 1 \
    -f -o $OUT/build_list python scripts/rebuild_cost.py > $OUT/ncu_list.log 2>&1
ls -la $OUT

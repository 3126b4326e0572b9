#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
DEMB200_TILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_force -f -o $OUT/tile \
    python scripts/profile_kernels.py --rebuilds 0 --steady 1 --settle 1500 > $OUT/ncu_tile.log 2>&1
tail -3 $OUT/ncu_tile.log

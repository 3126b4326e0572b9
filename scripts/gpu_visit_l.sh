#!/bin/bash
OUT=gpurun_out/r01l
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -4 $OUT/pytest_gpu.log
bash scripts/gpu_variants.sh r01l chrono_b200/libchrono_b200_dem.so build/variants/*.so

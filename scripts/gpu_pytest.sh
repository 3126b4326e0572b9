#!/bin/bash
# run a pytest selection on the GPU box: bash scripts/gpu_pytest.sh <tag> <pytest args...>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
( timeout 1500 python -m pytest "$@" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log )
tail -60 $OUT/pytest.log

#!/bin/bash
# A/B of library variants (bench, short) + the single-step parity tests with the LAST variant
TAG=$1; shift
bash scripts/gpu_variants.sh $TAG "$@"
last="${@: -1}"
( DEMB200_LIB=$PWD/$last timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/$TAG/pytest_parity.log 2>&1; echo "exit $?" >> gpurun_out/$TAG/pytest_parity.log )
tail -3 gpurun_out/$TAG/pytest_parity.log

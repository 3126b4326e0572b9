#!/bin/bash
# quick GPU visit: parity tests + smoke + a short bench.  usage: bash scripts/gpu_quick.sh <tag> [pytest args]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ${@:2} > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -30 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-steps 0 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json; tail -5 $OUT/bench.err

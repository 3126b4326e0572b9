#!/bin/bash
# Last visit of round 1: default (flat, profilable) step graph vs the opt-in conditional graph; tests; launch list.
TAG=${1:-r01final3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -3 $OUT/pytest_gpu.log
( timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json
DEMB200_COND_GRAPH=1 timeout 300 python bench.py --steps 10 --warmup 3 --cpu-steps 0 > $OUT/bench_condgraph.json 2> $OUT/bench_condgraph.err
cat $OUT/bench_condgraph.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 11700 -c 180 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 11 --warmup 3 --substeps 100 --cpu-steps 0 > $OUT/ncu_launch_bench.log 2>&1
wc -l $OUT/launches.csv

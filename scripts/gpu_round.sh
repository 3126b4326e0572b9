#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (ours + reference arm), steady-state ncu capture + launch list.
# usage (from repo root, under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -5 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json
bash scripts/gpu_ncu3.sh $TAG

#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the force kernel.
# usage (from repo root, under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -5 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
# launch list (serialised, cold cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --substeps 4 --cpu-steps 0 > $OUT/ncu_launch_bench.log 2>&1
# full capture of the dominant kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_force -s 20 -c 2 \
    -f -o $OUT/force python bench.py --steps 1 --warmup 3 --substeps 4 --cpu-steps 0 > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT

#!/bin/bash
# r02 GPU visit B: all GPU tests (new wall / contact-info tests included), default bench line with the equilibrium bed,
# ncu --set full of every kernel of a rebuilding and of a steady step.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -15 $OUT/pytest_gpu.log
( time timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ) 2>> $OUT/bench.err
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $OUT/kernels \
    python scripts/profile_kernels.py --rebuilds 1 --steady 1 > $OUT/ncu_kernels.log 2>&1
tail -3 $OUT/ncu_kernels.log
ls -la $OUT

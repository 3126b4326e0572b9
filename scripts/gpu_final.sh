#!/bin/bash
# final visit of a round: all GPU tests, smoke, the default bench line + the reference arm, ncu launch list of a short bench run
# (flat step graph: ncu cannot look into a conditional node), ncu --set full of every kernel (rebuilding + steady step)
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -4 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2>> $OUT/bench.err
cut -c1-400 $OUT/bench.json; tail -4 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cut -c1-300 $OUT/bench_ref.json
DEMB200_COND_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --substeps 10 --settle 2000 --cpu-steps 0 --no-incumbent --weak-base 0 --no-flowing > $OUT/bench_under_ncu.log 2>&1
tail -1 $OUT/launches.csv | cut -c1-200
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $OUT/kernels \
    python scripts/profile_kernels.py --rebuilds 1 --steady 1 > $OUT/ncu_kernels.log 2>&1
tail -2 $OUT/ncu_kernels.log

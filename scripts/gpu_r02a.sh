#!/bin/bash
# r02 GPU visit A: parity tests + smoke, the new default bench line (settled bed, flowing record, incumbent, 32 M base),
# configs[0], the incumbent alone.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -5 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
( timeout 300 baseline/_ref/incumbent_dem --spheres 1000000 --warmup 100 --steps 200 > $OUT/incumbent.json 2> $OUT/incumbent.err; echo "rc $?" >> $OUT/incumbent.err )
cat $OUT/incumbent.json; tail -3 $OUT/incumbent.err
( time timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ) 2>> $OUT/bench.err
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --config 0 > $OUT/bench_cfg0.json 2> $OUT/bench_cfg0.err
cat $OUT/bench_cfg0.json; tail -3 $OUT/bench_cfg0.err

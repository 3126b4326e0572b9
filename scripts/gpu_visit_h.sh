#!/bin/bash
# GPU visit r01h (2 GPUs): slab run bit-identical to the single-GPU run (with and without the lagged vote), bench at N=2.
OUT=gpurun_out/r01h
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29511 tests/mgpu_check.py --spheres 20000 --steps 300 > $OUT/mgpu_check_lag0.log 2>&1; echo "exit $?" >> $OUT/mgpu_check_lag0.log )
grep -E "owned|PASSED|exit|Error|error" $OUT/mgpu_check_lag0.log | tail -5
( timeout 600 $TR --master-port 29512 tests/mgpu_check.py --spheres 20000 --steps 300 --lag 1 > $OUT/mgpu_check_lag1.log 2>&1; echo "exit $?" >> $OUT/mgpu_check_lag1.log )
grep -E "owned|PASSED|exit|Error|error" $OUT/mgpu_check_lag1.log | tail -5
( timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "exit $?" >> $OUT/bench_n2.err )
cat $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --cpu-steps 0 > $OUT/bench_n1.json 2> $OUT/bench_n1.err )
cat $OUT/bench_n1.json

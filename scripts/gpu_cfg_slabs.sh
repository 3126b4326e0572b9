#!/bin/bash
# usage: NR=<ranks> CFG=<config> bash scripts/gpu_cfg_slabs.sh <tag>   -- a named config in slab mode
TAG=${1:-r02slab}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NR --master-addr 127.0.0.1 --master-port 29511"
( time timeout ${TMO:-500} $TR bench.py --gpus $NR --config $CFG --steps ${STEPS:-10} --warmup 3 $EXTRA > $OUT/bench_cfg${CFG}_n$NR.json 2> $OUT/bench_cfg${CFG}_n$NR.err; echo "exit $?" >> $OUT/bench_cfg${CFG}_n$NR.err ) 2>> $OUT/bench_cfg${CFG}_n$NR.err
grep "^{" $OUT/bench_cfg${CFG}_n$NR.json | cut -c1-4000; grep -E "Error|error|exit|real" $OUT/bench_cfg${CFG}_n$NR.err | tail -6

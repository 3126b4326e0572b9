#!/bin/bash
TAG=${1:-r02n2cfg4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NR:-2} --master-addr 127.0.0.1 --master-port 29511"
( time timeout ${TMO:-800} $TR bench.py --gpus ${NR:-2} --steps ${STEPS:-10} --warmup 3 $EXTRA > $OUT/bench.json 2> $OUT/bench.err; echo "exit $?" >> $OUT/bench.err ) 2>> $OUT/bench.err
grep "^{" $OUT/bench.json | cut -c1-6000; tail -6 $OUT/bench.err; nvidia-smi --query-gpu=memory.used --format=csv | head -3

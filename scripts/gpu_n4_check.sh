#!/bin/bash
TAG=${1:-r02n4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
NR=${NR:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NR --master-addr 127.0.0.1 --master-port 29511"
( timeout 150 $TR tests/mgpu_check.py --p2p --spheres ${NS:-100000} --steps 300 > $OUT/mgpu_check.log 2>&1; echo "exit $?" >> $OUT/mgpu_check.log )
grep -E "owned per rank|PASSED|DemError|exit" $OUT/mgpu_check.log | tail -8 | cut -c1-600

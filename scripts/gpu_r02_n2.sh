#!/bin/bash
# 2 x B200: slab run vs single GPU bit for bit (plain and with a moving mesh), then the 1 M / GPU weak-scaling line (configs[1] physics)
TAG=${1:-r02n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( timeout 180 $TR tests/mgpu_check.py --p2p --spheres 40000 --steps 400 > $OUT/mgpu_check.log 2>&1; echo "exit $?" >> $OUT/mgpu_check.log )
grep -E "owned per rank|PASSED|Error|error|exit" $OUT/mgpu_check.log | tail -5
( timeout 180 $TR tests/mgpu_check.py --p2p --mesh --spheres 20000 --steps 300 > $OUT/mgpu_check_mesh.log 2>&1; echo "exit $?" >> $OUT/mgpu_check_mesh.log )
grep -E "owned per rank|PASSED|mesh force|Error|error|exit" $OUT/mgpu_check_mesh.log | tail -5
( timeout 300 $TR bench.py --gpus 2 --config 1 --steps 10 --warmup 3 > $OUT/bench_n2_cfg1.json 2> $OUT/bench_n2_cfg1.err; echo "exit $?" >> $OUT/bench_n2_cfg1.err )
cat $OUT/bench_n2_cfg1.json | cut -c1-3500; tail -3 $OUT/bench_n2_cfg1.err

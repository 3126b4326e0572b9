#!/bin/bash
# GPU visit r01i (2 GPUs): P2P halo path: bit-identical check, bench N=2 with P2P and with NCCL.
OUT=gpurun_out/r01i
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/mgpu_check.py --spheres 20000 --steps 300 --p2p > $OUT/mgpu_check_p2p.log 2>&1; echo "exit $?" >> $OUT/mgpu_check_p2p.log )
grep -E "owned|PASSED|exit|Error|error" $OUT/mgpu_check_p2p.log | tail -8
( timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2_p2p.json 2> $OUT/bench_n2_p2p.err; echo "exit $?" >> $OUT/bench_n2_p2p.err )
cat $OUT/bench_n2_p2p.json; tail -3 $OUT/bench_n2_p2p.err
( timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --nccl-halo > $OUT/bench_n2_nccl.json 2> $OUT/bench_n2_nccl.err; echo "exit $?" >> $OUT/bench_n2_nccl.err )
grep "^{" $OUT/bench_n2_nccl.json | cut -c1-300

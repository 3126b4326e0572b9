#!/bin/bash
# usage: bash scripts/gpu_visit_n4.sh <ngpus> <tag>
N=${1:-4}; TAG=${2:-r01p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/mgpu_check.py --spheres 40000 --steps 300 --p2p > $OUT/mgpu_check_p2p_n$N.log 2>&1; echo "exit $?" >> $OUT/mgpu_check_p2p_n$N.log )
grep -E "owned|PASSED|exit|Error|error" $OUT/mgpu_check_p2p_n$N.log | tail -6
( timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "exit $?" >> $OUT/bench_n$N.err )
grep "^{" $OUT/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('N', d['n_gpus'], 'value %.3e'%d['value'], 'ms/step', round(d['ms_per_step'],2), 'e2e %.3e'%d['e2e']['value'], 'halo', d['halo']['nvlink_GBps_busiest_rank'], 'rebuild ms', d['halo']['rebuild_ms_total_in_timed_region'], 'us/ts outside', 1e3*d['halo']['ms_per_timestep_outside_rebuilds'])"
tail -3 $OUT/bench_n$N.err

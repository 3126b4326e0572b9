#!/bin/bash
OUT=gpurun_out/r01j
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for cfg in "--slab-lag 2" "--slab-lag 3" "--slab-lag 4" "--slab-lag 2 --skin 0.35" "--slab-lag 3 --skin 0.35" "--slab-lag 3 --skin 0.5"; do
  i=$((i+1))
  timeout 300 $TR --master-port $((29520+i)) bench.py --gpus 2 --steps 5 --warmup 3 $cfg 2>/dev/null | grep "^{" > $OUT/sweep_$i.json
  python -c "
import json,sys
d=json.load(open('$OUT/sweep_$i.json'))
print('$cfg', 'ms/step', round(d['ms_per_step'],2), 'rebuilds', d['halo']['rebuilds'], 'rebuild_ms', round(d['halo']['rebuild_ms_total_in_timed_region'][0],1), 'us/ts outside', round(1e3*d['halo']['ms_per_timestep_outside_rebuilds'],1), 'cbar', round(d['contacts_per_sphere'],2))"
done

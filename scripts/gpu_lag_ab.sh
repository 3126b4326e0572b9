#!/bin/bash
TAG=${1:-r02lag}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for lag in 3; do
  timeout 300 $TR bench.py --gpus 2 --config 1 --steps 5 --warmup 3 --settle 2000 --no-parity --slab-lag $lag 2>/dev/null | grep "^{" > $OUT/n2_lag$lag.json
  python - <<PY
import json
b=json.load(open("$OUT/n2_lag$lag.json"))
print("lag $lag  N=2 %.3e (%.4f ms/ts) wall %.4f enqueue %.4f"%(b["value"], b["ms_per_step"]/100, b["wall_s_timed_region"], b["host_enqueue_s"]))
PY
done

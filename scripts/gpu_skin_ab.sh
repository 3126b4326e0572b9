#!/bin/bash
TAG=${1:-r02skin}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for sk in 0.25 0.35; do
  DEMB200_COND_GRAPH=0 timeout 200 python bench.py --steps 5 --warmup 3 --cpu-steps 0 --no-incumbent --weak-base 0 --no-flowing --settle 2000 --skin $sk > $OUT/n1_skin$sk.json 2>/dev/null
  timeout 300 $TR bench.py --gpus 2 --config 1 --steps 5 --warmup 3 --settle 2000 --no-parity --skin $sk 2>/dev/null | grep "^{" > $OUT/n2_skin$sk.json
  python - <<PY
import json
a=json.load(open("$OUT/n1_skin$sk.json")); b=json.load(open("$OUT/n2_skin$sk.json"))
print("skin $sk  N=1 %.3e (%.4f ms/ts, force %.1f us)   N=2 %.3e (%.4f ms/ts)  eff %.3f"%(a["value"], a["ms_per_step"]/100, 1000*a["kernel_ms_per_timestep"]["k_force_integrate"], b["value"], b["ms_per_step"]/100, b["value"]/2/a["value"]))
PY
done

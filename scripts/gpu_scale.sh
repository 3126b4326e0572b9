#!/bin/bash
# single-GPU size sweep of the bench workload (memory sizing for config[4]: 8-32 M spheres per GPU)
OUT=gpurun_out/r01s
mkdir -p $OUT
for n in 4000000 16000000 32000000; do
  timeout 900 python bench.py --spheres $n --steps 2 --warmup 3 --substeps 20 --cpu-steps 0 > $OUT/bench_$n.json 2> $OUT/bench_$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$n.json"))
    print("$n", "value %.3e"%d["value"], "ms/timestep %.3f"%(d["ms_per_step"]/20), "cbar %.2f"%d["contacts_per_sphere"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e:
    print("$n failed", e); print(open("$OUT/bench_$n.err").read()[-800:])
PY
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done

#!/bin/bash
# bench several (library, DEMB200_CARVEOUT) pairs: bash scripts/gpu_variants2.sh <tag> lib1.so[:carveout] ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for spec in "$@"; do
  lib=${spec%%:*}; co=""; [[ "$spec" == *:* ]] && co=${spec##*:}
  name=$(basename $lib .so)${co:+_co$co}
  DEMB200_LIB=$PWD/$lib ${co:+env DEMB200_CARVEOUT=$co} timeout 600 python bench.py --steps 5 --warmup 3 --cpu-steps 0 --no-incumbent --weak-base 0 --no-flowing --settle 2000 > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json"))
    print("$name", "value %.3e ms/step %.2f force_us %.1f frac %.3f cbar %.2f"%(d["value"], d["ms_per_step"], 1000*d["kernel_ms_per_timestep"]["k_force_integrate"], d["roofline"]["frac"], d["contacts_per_sphere"]))
except Exception as e:
    print("$name", "FAILED", e)
PY
done

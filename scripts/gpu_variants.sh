#!/bin/bash
# bench several prebuilt library variants: bash scripts/gpu_variants.sh <tag> lib1.so lib2.so ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  DEMB200_LIB=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --cpu-steps 0 --no-incumbent --weak-base 0 --no-flowing --settle 2000 > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$name.json"))
print("$name", "value %.3e ms/step %.2f force_us %.1f frac %.3f cbar %.2f"%(d["value"], d["ms_per_step"], 1000*d["kernel_ms_per_timestep"]["k_force_integrate"], d["roofline"]["frac"], d["contacts_per_sphere"]))
PY
done

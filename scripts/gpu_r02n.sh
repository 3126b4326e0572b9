#!/bin/bash
# all GPU tests, short bench (e2e with the no-rebuild round trip), per-kernel cost of a rebuilding step
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-steps 0 --no-incumbent --weak-base 0 > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.4e e2e %.4e flowing %.4e force_us %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["flowing"]["value"], 1000*d["kernel_ms_per_timestep"]["k_force_integrate"], d["roofline"]["frac"]), d["roofline"]["traffic"])
PY
timeout 300 python scripts/rebuild_cost.py > $OUT/rebuild_cost.log 2>&1
cat $OUT/rebuild_cost.log | tail -4

#!/bin/bash
# confirmation visit after a kernel change: all GPU tests, smoke, the default bench line
TAG=${1:-r02confirm}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -4 $OUT/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log )
tail -2 $OUT/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2>> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.4e e2e %.4e flowing %.4e force_us %.1f frac %.3f whole %.3f weak_base %.4e" % (d["value"], d["e2e"]["value"], d["flowing"]["value"], 1000*d["kernel_ms_per_timestep"]["k_force_integrate"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["weak_base"]["value"]))
PY
tail -3 $OUT/bench.err

#!/bin/bash
# tile kernel bring-up: sanitizer on a small case, tile-vs-plain bitwise test, oracle parity suite with the tile path, bench A/B
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export DEMB200_LIB=$PWD/build/lib_tile.so
( DEMB200_TILE=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "force_models_three_steps or ragged" > $OUT/sanitizer.log 2>&1; echo "sanitizer exit $?" >> $OUT/sanitizer.log )
tail -8 $OUT/sanitizer.log
( timeout 600 python -m pytest tests/test_gpu_tile.py -x -q > $OUT/pytest_tile.log 2>&1; echo "exit $?" >> $OUT/pytest_tile.log )
tail -12 $OUT/pytest_tile.log
( DEMB200_TILE=1 timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_tile1.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu_tile1.log )
tail -6 $OUT/pytest_gpu_tile1.log
for t in 0 1; do
  DEMB200_TILE=$t timeout 600 python bench.py --steps 5 --warmup 3 --cpu-steps 0 --no-incumbent --weak-base 0 --settle 2000 > $OUT/bench_tile$t.json 2> $OUT/bench_tile$t.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_tile$t.json"))
print("tile=$t", "value %.3e ms/step %.2f frac %.3f cbar %.2f flowing %.3e"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["contacts_per_sphere"], d["flowing"]["value"]), {k: round(1000*v,1) for k,v in d["kernel_ms_per_timestep"].items()})
PY
done

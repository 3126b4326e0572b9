#!/bin/bash
OUT=gpurun_out/r01k
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-steps 0 > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json'))
print('value', d['value']/1e9, 'ms/step', d['ms_per_step'], 'force ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'rebuilds', d['neighbor_list_rebuilds'])
print(d['kernel_ms_per_timestep'])"

#!/bin/bash
# Named sizes of BASELINE configs[2] / configs[3] on ONE GPU (their multi-GPU form is the slab mode of bench.py --gpus N).
TAG=${1:-r01sizes}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python bench.py --spheres 4000000 --polydisperse 0.8,1.2 --mu-roll 0.05 --cpu-steps 0 --steps 5 --warmup 3 \
    > $OUT/bench_cfg2_4M_poly_roll.json 2> $OUT/bench_cfg2.err
cat $OUT/bench_cfg2_4M_poly_roll.json | cut -c1-600
timeout 240 python scripts/bench_drum.py --spheres 8000000 --segments 160 --axial 159 --steps 200 --warm 100 \
    > $OUT/drum_8M.json 2> $OUT/drum_8M.err
cat $OUT/drum_8M.json; tail -3 $OUT/drum_8M.err

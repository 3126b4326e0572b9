#!/bin/bash
# BASELINE configs[2] (4 M polydisperse + rolling) and configs[3] (8 M drum, bed on the wall) on ONE GPU
TAG=${1:-r02cfg23}
OUT=gpurun_out/$TAG; mkdir -p $OUT
( time timeout 400 python bench.py --config 2 --steps 10 --warmup 3 --cpu-steps 2 > $OUT/bench_cfg2_n1.json 2> $OUT/bench_cfg2_n1.err ) 2>> $OUT/bench_cfg2_n1.err
cut -c1-2500 $OUT/bench_cfg2_n1.json; tail -4 $OUT/bench_cfg2_n1.err
( time timeout 600 python bench.py --config 3 --steps 5 --warmup 2 --settle 3000 > $OUT/bench_cfg3_n1.json 2> $OUT/bench_cfg3_n1.err ) 2>> $OUT/bench_cfg3_n1.err
cut -c1-2500 $OUT/bench_cfg3_n1.json; tail -4 $OUT/bench_cfg3_n1.err

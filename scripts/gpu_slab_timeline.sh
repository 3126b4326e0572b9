#!/bin/bash
# per-kernel launch list of slab steps (every rank under its own ncu, durations only: one pass, no replay)
# usage: bash scripts/gpu_slab_timeline.sh <tag> <nranks> [spheres per gpu]
TAG=${1:-r02tl}; NR=${2:-2}; NS=${3:-1000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $NR --master-addr 127.0.0.1 --master-port 29533 \
  bash -c "DEMB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 48000 -c 450 --csv --log-file $OUT/launches_rank\$RANK.csv \
  python bench.py --gpus $NR --config 1 --spheres $NS --steps 3 --warmup 2 --settle 3000 --no-parity > $OUT/bench_rank\$RANK.out 2> $OUT/bench_rank\$RANK.err"
echo "rc $?"
ls -la $OUT
python - <<PY
import csv, collections, glob
for f in sorted(glob.glob("$OUT/launches_rank*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>5 and r[0].isdigit()]
    if not rows: print(f, "no rows"); continue
    # columns: ID, PID, process, host, kernel, context, stream, block, grid, device, cc, section, metric, unit, value
    agg=collections.OrderedDict()
    for r in rows:
        k=r[4].split("(")[0]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1].replace(",",""))
    print(f, len(rows), "launches")
    for k,(n,t) in agg.items(): print("   %-60s n=%4d mean %.1f us"%(k[:60], n, t/n/ (1000.0 if 'ns' in rows[0][-2] else 1.0)))
PY

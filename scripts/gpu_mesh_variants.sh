for lib in build/variants/*.so; do name=$(basename $lib .so); echo -n "$name "; DEMB200_LIB=$PWD/$lib python scripts/bench_drum.py --warm 3000 --steps 500 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['sphere_steps_per_s']/1e9, d['ms_per_timestep'], d['rebuilds_in_timed_region'])"; done

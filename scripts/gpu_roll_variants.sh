for lib in build/variants/*.so; do name=$(basename $lib .so); DEMB200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --cpu-steps 0 --polydisperse 0.8,1.2 --mu-roll 0.05 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$name', d['value']/1e9, d['ms_per_step'], d['kernel_ms_per_timestep']['k_force_integrate'])"; done

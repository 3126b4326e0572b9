import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench, dem_common as common
from chrono_b200 import dem
sc = bench.build_scene(1000000)
for skin in (-1.0, 0.0):
    g = common.make_gpu(sc, dt=1e-4, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP, verlet_skin=skin)
    g.step(50)
    ms = g.step_timed(100)
    st = g.stats()
    print("skin", skin, "ms/step", ms / 100, "rebuilds", st["rebuilds"], flush=True)
    prof = g.step_profile(20)
    print({k: round(v / 20, 4) for k, v in prof.items()})
    g.close()

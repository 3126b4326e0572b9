"""Per-kernel device time of a step that rebuilds the neighbour lists (Verlet skin 0: every step rebuilds) next to the default
(settled 1 M-sphere bed of the bench):  python scripts/rebuild_cost.py [spheres]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chrono_b200 import scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sc = bench.build_scene(1, n)
for skin in (-1.0, 0.0):
    kw = dict(bench.physics(1))
    kw["verlet_skin"] = skin
    g = scenes.make_gpu(sc, **kw)
    g.step(200)
    ms = g.step_timed(100)
    st = g.stats()
    print("skin", skin, "ms/step", ms / 100, "rebuilds", st["rebuilds"], flush=True)
    prof = g.step_profile(20)
    print({k: round(v / 20, 4) for k, v in prof.items()}, flush=True)
    g.close()

"""Run under ncu with --profile-from-start off: the 1 M-sphere settled bed of the default bench line, then, inside
cudaProfilerStart/Stop, `--rebuilds` time steps that rebuild the neighbour lists (requested through dem_b200_request_rebuild, so the Verlet skin and the list lengths are the ones of a normal run) and `--steady` steps that do not.

  ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/TAG/kernels \
      python scripts/profile_kernels.py --rebuilds 2 --steady 2
"""
import argparse
import os
import sys

import numpy as np
import torch

os.environ.setdefault("DEMB200_COND_GRAPH", "0")  # ncu cannot see the kernel nodes of a graph with a conditional node

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chrono_b200 import dem, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=1000000)
    ap.add_argument("--settle", type=int, default=3000)
    ap.add_argument("--rebuilds", type=int, default=2)
    ap.add_argument("--steady", type=int, default=2)
    ap.add_argument("--config", type=int, default=1)
    args = ap.parse_args()
    scene = bench.build_scene(args.config, args.spheres)
    g = scenes.make_gpu(scene, **bench.physics(args.config))
    g.step(args.settle)
    g.request_rebuild()
    g.step(1)  # one rebuild outside the capture: the storage order is the settled one from here on
    g.step(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.rebuilds):
        g.request_rebuild()  # the Verlet skin and the list lengths are those of a normal run
        g.step(1)
    for _ in range(args.steady):
        g.step(1)
    g.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("contacts per sphere", g.reduce(dem.RED_NUM_CONTACTS) / scene["n"], "stats", g.stats())


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE configs[3]-style workload on one GPU: spheres in a rotating drum made of a triangle mesh (axis y,
omega = 1 rad/s through ApplyMeshMotion before EVERY step), Hertz-Mindlin MultiStep.  Prints one JSON line.
Not a bench.py line (bench.py measures configs[1]); this is the mesh-path throughput check."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=2000000)
    ap.add_argument("--segments", type=int, default=100, help="drum segments around")
    ap.add_argument("--axial", type=int, default=99, help="subdivisions along the axis (2 triangles per patch, + 2 cap fans)")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warm", type=int, default=100)
    args = ap.parse_args()
    from chrono_b200 import dem, scenes
    import dem_common as common
    R = 0.02
    n = args.spheres
    # half-filled drum: volume per sphere (HCP at 2R) = (2R)^3 / sqrt(2); drum length = diameter
    vol = n * (2 * R) ** 3 / math.sqrt(2) / 0.5 * 1.15
    Rd = (vol / (2 * math.pi)) ** (1.0 / 3.0)
    Ld = 2 * Rd
    tri = scenes.cylinder_drum_mesh(Rd, Ld, args.segments, n_ax=args.axial)
    pts = scenes.hcp_points((-Rd, -Ld / 2 + 1.5 * R, -Rd), (Rd, Ld / 2 - 1.5 * R, 0.0), 2.0 * R)
    pts = pts[np.hypot(pts[:, 0], pts[:, 2]) < Rd - 1.5 * R]
    rng = np.random.default_rng(1)
    pts = pts[:n] + rng.uniform(-0.005 * R, 0.005 * R, size=(min(n, len(pts)), 3))
    n = len(pts)
    sc = dict(pos=pts, radius=np.full(n, R), walls=[], bins=(10, 10, 10), n=n,
              meshes=[dict(tri=tri, pos=np.zeros(3), rot=np.array([1.0, 0, 0, 0]), vel=np.zeros(3), omega=np.array([0, 1.0, 0]), mass=1e3)])
    g = common.make_gpu(sc, dt=1e-4, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP, history_slots=20)

    def advance(k0, k1):
        for it in range(k0, k1):
            q = scenes.quat_from_axis_angle((0, 1, 0), 1.0 * it * 1e-4)
            g.set_mesh_motion(0, None, q, None, (0, 1.0, 0))
            g.step(1, sync=False)
        g.sync()

    advance(0, args.warm)  # warm-up: first contacts, first rebuilds, graph capture
    r0 = g.stats()["rebuilds"]
    t0 = time.perf_counter()
    advance(args.warm, args.warm + args.steps)
    dt = time.perf_counter() - t0
    st = g.stats()
    f, tq = g.mesh_wrench(0)
    p, v, _ = g.state()
    line = {"workload": "rotating drum, triangle mesh, ApplyMeshMotion every step", "spheres": n, "triangles": int(len(tri)),
            "timesteps": args.steps, "sphere_steps_per_s": n * args.steps / dt, "ms_per_timestep": 1e3 * dt / args.steps,
            "rebuilds_in_timed_region": st["rebuilds"] - r0, "contacts_per_sphere": g.reduce(dem.RED_NUM_CONTACTS) / n,
            "mesh_force": f.tolist(), "weight_of_bed": float(common.sphere_mass(R) * n * 9.81),
            "all_inside_drum": bool((np.hypot(p[:, 0], p[:, 2]) < Rd).all()), "timing": "host wall clock around the step loop, synchronised"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()

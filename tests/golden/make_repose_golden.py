"""Generates tests/golden/repose_oracle.json: the CPU oracle's angle of repose over an ENSEMBLE of piles.

Why an ensemble: once GPU and oracle trajectories have diverged (different summation orders amplify 1e-16 differences), two
runs of the same pile differ like two different seeds do.  Measured here with the oracle alone: the angle of a 1 200-sphere
pile scatters by 7 % (least-squares slope of the surface) or 1.9 % (moment estimator below) from seed to seed, and a
5 200-sphere pile scatters just as much (avalanches, not counting noise).  A 1 % statement (BASELINE.json north_star) is
therefore a statement about ensemble means: 128 piles bring the standard error of a mean to 0.22 %.

The oracle takes ~30 s per pile: the ensemble is computed once, here, and committed; tests/test_gpu_longrun.py runs the same
128 scenes on the GPU (two minutes) and compares the means.  Re-run: python tests/golden/make_repose_golden.py [nproc]
"""
import json
import math
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEEDS = list(range(100, 228))
STEPS = 14000


def one(seed):
    import dem_common as common
    import test_gpu_longrun as T
    from oracle import pyoracle as po
    sc = T.repose_scene(seed)
    o = common.make_oracle(sc, num_threads=2, **T.repose_physics(po))
    assert o.step(STEPS) == 0
    p, _, v, _ = o.state()
    f = o.first_sphere_body
    a, zbar, rho = T.moment_angle(p[f:], sc["radius"])
    return dict(seed=seed, n=int(sc["n"]), angle_deg=a, zbar=zbar, rho_rms=rho,
                v99=float(np.quantile(np.linalg.norm(v[f:], axis=1), 0.99)))


def main():
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    with ProcessPoolExecutor(nproc) as ex:
        rows = list(ex.map(one, SEEDS))
    a = np.array([r["angle_deg"] for r in rows])
    out = dict(what="CPU oracle, polydisperse pile released on a plane (tests/test_gpu_longrun.py::repose_scene), %d steps of 1e-4 s" % STEPS,
               estimator="atan(4 zbar / sqrt(10/3 <rho^2>)): angle of the cone with the pile's mass moments", steps=STEPS,
               mean_angle_deg=float(a.mean()), std_angle_deg=float(a.std(ddof=1)), sem_angle_deg=float(a.std(ddof=1) / math.sqrt(len(a))),
               mean_zbar=float(np.mean([r["zbar"] for r in rows])), mean_rho_rms=float(np.mean([r["rho_rms"] for r in rows])), piles=rows)
    with open(os.path.join(HERE, "repose_oracle.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("mean angle %.3f deg, std %.3f, sem %.3f over %d piles" % (out["mean_angle_deg"], out["std_angle_deg"], out["sem_angle_deg"], len(a)))


if __name__ == "__main__":
    main()

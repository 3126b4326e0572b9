"""Generate tests/golden/ref_vectors.npz from the REFERENCE's own object code (oracle/_ref/libchrono_ref.so).

Run in the build container (where /root/reference is mounted):   python tests/golden/make_golden.py
The fixture travels to the GPU box, where the reference does not exist; tests/test_oracle_fixtures.py replays it
against the oracle (bit-exact) and tests/test_gpu_parity.py against the CUDA kernels.
Contents (all seeded, small):
  prim_*      random sphere_sphere / box_sphere / triangle_sphere inputs and the reference outputs
  force_*     random single contacts through function_CalcContactForces (4 models x matprops x 3 tangential modes)
  scene_*     a 600-sphere five-wall scene: reference grid, bin CSR, candidate pairs, contacts
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402
from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def pack(r):
    """contact dict|None -> 12 doubles (hit, norm3, depth, pt1_3, pt2_3, erad)"""
    if r is None:
        return np.zeros(12)
    return np.concatenate([[1.0], r["norm"], [r["depth"]], r["pt1"], r["pt2"], [r["erad"]]])


def main():
    po.build(ref=True)
    R = po.ref()
    rng = np.random.default_rng(20261017)
    out = {}
    # ---- primitives
    n = 400
    ss_in, ss_out, bs_in, bs_out, ts_in, ts_out = [], [], [], [], [], []
    for _ in range(n):
        p1, p2 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        r1, r2 = rng.uniform(0.2, 0.9, 2)
        sep = float(rng.choice([0.0, 0.05]))
        ss_in.append(np.concatenate([p1, [r1], p2, [r2, sep]]))
        ss_out.append(pack(R.prims.sphere_sphere(p1, r1, p2, r2, sep)))
        q, hd = rand_quat(rng), rng.uniform(0.2, 1.0, 3)
        bs_in.append(np.concatenate([p1, q, hd, p2 * 2, [r2, sep]]))
        bs_out.append(pack(R.prims.box_sphere(p1, q, hd, p2 * 2, r2, sep)))
        A, B, Cc = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        ts_in.append(np.concatenate([A, B, Cc, p2, [r2, sep]]))
        ts_out.append(pack(R.prims.triangle_sphere(A, B, Cc, p2, r2, sep)))
    out.update(prim_ss_in=np.array(ss_in), prim_ss_out=np.array(ss_out), prim_bs_in=np.array(bs_in),
               prim_bs_out=np.array(bs_out), prim_ts_in=np.array(ts_in), prim_ts_out=np.array(ts_out))
    # ---- single-contact forces
    fin, fout = [], []
    for model in (po.HERTZ, po.HOOKE, po.FLORES, po.PLAINCOULOMB):
        for mat_props in (1, 0):
            for tang in (po.TANG_NONE, po.TANG_ONESTEP, po.TANG_MULTISTEP):
                for _ in range(12):
                    adh = int(rng.integers(0, 3))
                    dt = float(rng.choice([1e-3, 1e-4]))
                    s = po.make_settings(force_model=model, adhesion_model=adh, tangential_mode=tang,
                                         use_mat_props=bool(mat_props), dt=dt)
                    mat = [float(rng.uniform(1e5, 1e7)), 0.3, float(rng.uniform(0, 0.8)), float(rng.choice([0.0, 0.05])),
                           float(rng.choice([0.0, 0.02])), float(rng.uniform(0.05, 0.95)), float(rng.choice([0.0, 0.3])),
                           0.1, 0.2, 2e5, 1e5, 40.0, 20.0]
                    m = po.OrcMaterial(*mat)
                    comp = po.composite(m, m)
                    b1, b2 = (0, 1) if rng.random() < 0.5 else (1, 0)
                    r = rng.uniform(0.01, 0.05, 2)
                    mass = common.sphere_mass(r)
                    pos = np.zeros((2, 3))
                    pos[0] = rng.uniform(-1, 1, 3)
                    d = rng.normal(size=3)
                    d /= np.linalg.norm(d)
                    depth = -float(rng.uniform(1e-6, 2e-3))
                    pos[1] = pos[0] + d * (r[0] + r[1] + depth)
                    rot = np.array([rand_quat(rng), rand_quat(rng)])
                    vel = rng.normal(size=(2, 6)) * np.array([0.3, 0.3, 0.3, 5, 5, 5])
                    nrm = (pos[b2] - pos[b1]) / np.linalg.norm(pos[b2] - pos[b1])
                    pt1, pt2 = pos[b1] + nrm * r[b1], pos[b2] - nrm * r[b2]
                    erad = r[0] * r[1] / (r[0] + r[1])
                    has_h = tang == po.TANG_MULTISTEP and rng.random() < 0.7
                    h = dict(disp=rng.normal(size=3) * 1e-5, dur=float(rng.uniform(0, 0.05)),
                             relvel=float(rng.uniform(0, 2))) if has_h else None
                    F, T1, T2, ho = po.contact_force("ref", s, comp, b1, b2, mass, pos, rot, vel, pt1, pt2, nrm, depth,
                                                     erad, h)
                    fin.append(np.concatenate([[model, adh, tang, mat_props, dt], mat, [b1, b2], mass, pos.ravel(),
                                               rot.ravel(), vel.ravel(), pt1, pt2, nrm, [depth, erad],
                                               [1.0 if h else 0.0], h["disp"] if h else np.zeros(3),
                                               [h["dur"] if h else 0.0, h["relvel"] if h else 0.0]]))
                    fout.append(np.concatenate([F, T1, T2, [1.0 if ho else 0.0], ho["disp"] if ho else np.zeros(3),
                                                [ho["dur"] if ho else 0.0, ho["relvel"] if ho else 0.0]]))
    out.update(force_in=np.array(fin), force_out=np.array(fout))
    # ---- one scene through the reference broadphase + narrowphase
    scene = scenes.settling_scene(600, sep_factor=1.98, seed=7)
    o = common.make_oracle(scene)
    mn, mx = o.generate_aabb()
    nW, ns = len(scene["walls"]), len(scene["walls"]) + scene["n"]
    types = np.array([po.SHAPE_BOX] * nW + [po.SHAPE_SPHERE] * scene["n"], dtype=np.int32)
    bodies = np.array([0] * nW + list(range(1, scene["n"] + 1)), dtype=np.int32)
    lpos, dims = np.zeros((ns, 3)), np.zeros((ns, 3))
    for k, (p, h) in enumerate(scene["walls"]):
        lpos[k], dims[k] = p, h
    dims[nW:, 0] = scene["radius"]
    pos, rot, _, _ = o.state()
    active = np.ones(len(pos), dtype=np.int8)
    active[0] = 0
    r = R.collision(types, bodies, lpos, np.tile([1.0, 0, 0, 0], (ns, 1)), dims, np.zeros((ns, 9)), pos, rot, active,
                    np.ones(len(pos), dtype=np.int8), mn, mx, scene["bins"])
    out.update(scene_n=np.array([600]), scene_seed=np.array([7]), scene_origin=r["origin"], scene_bin_size=r["bin_size"],
               scene_inv_bin_size=r["inv_bin_size"], scene_bin_active=r["bin_active"],
               scene_bin_start=r["bin_start_index"], scene_bin_aabb=r["bin_aabb_number"], scene_pairs=r["pairs"],
               scene_ct_shape=r["contacts"]["shape_pair"], scene_ct_norm=r["contacts"]["normal"],
               scene_ct_depth=r["contacts"]["depth"], scene_ct_pt1=r["contacts"]["pt1"],
               scene_ct_pt2=r["contacts"]["pt2"], scene_ct_erad=r["contacts"]["erad"])
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

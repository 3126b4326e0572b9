"""Multi-GPU slab run vs a single-GPU run of the same scene, bit for bit.  Launch under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py [--spheres 20000] [--steps 300]

Every rank also runs the WHOLE scene on its own GPU with the plain engine and compares the spheres it owns at the
end (positions, velocities, angular velocities: np.array_equal).  Spheres get a horizontal drift so that some cross
the slab faces (migration with history) and the Verlet lists are rebuilt several times."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lag", type=int, default=0, help="1: read the rebuild vote one step late (no host sync per step)")
    ap.add_argument("--mesh", action="store_true", help="add a moving height-field mesh above the floor (sphere-facet history migrates too)")
    ap.add_argument("--p2p", action="store_true", help="halo through NVLink peer stores + device-side vote (no NCCL per step)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from chrono_b200 import dem, scenes, slab
    import dem_common as common
    n = args.spheres
    scene = scenes.settling_scene(n, sep_factor=1.99, seed=77)
    if args.mesh:
        L = scene["box_size"][0]
        R = float(scene["radius"].max())
        ncell = max(4, int(L / (1.4 * R)))
        bump = np.random.default_rng(11).uniform(0.0, 0.12 * R, size=(ncell + 1, ncell + 1))
        tri = scenes.heightfield_mesh(-L / 2, L / 2, -L / 2, L / 2, ncell, ncell, lambda X, Y: 0.03 * R + bump)
        scene["meshes"] = [dict(tri=tri, pos=np.zeros(3), rot=np.array([1.0, 0, 0, 0]), vel=np.array([0.05, 0.0, 0.0]),
                                omega=np.zeros(3), mass=5.0)]
    rng = np.random.default_rng(5)
    vel = rng.normal(size=(n, 3)) * 0.1
    vel[:, 0] += np.where(scene["pos"][:, 0] < 0, 0.6, -0.6)  # both halves drift towards (and across) the middle
    om = rng.normal(size=(n, 3)) * 2.0
    kw = dict(dt=1e-4, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP)

    # ---- reference: the whole scene on this GPU
    def mesh_pose(it):  # the mesh slides along x: ApplyMeshMotion before every step, on every rank
        return np.array([0.05 * it * 1e-4, 0.0, 0.0])

    ref = common.make_gpu(scene, vel=vel, omega=om, device=local, **kw)
    if args.mesh:
        wr = np.zeros(3)
        for it in range(args.steps):
            ref.set_mesh_motion(0, pos=mesh_pose(it))
            ref.step(1)
        wr = ref.mesh_wrench(0)[0]
    else:
        ref.step(args.steps)
    rp, rv, rw = ref.state()

    # ---- slab run
    bounds = slab.slab_bounds(scene["pos"][:, 0], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = np.nonzero((scene["pos"][:, 0] >= lo) & (scene["pos"][:, 0] < hi))[0]
    mat = common.settling_material()
    cfg = dem.config(device=local, bins=scene["bins"], mat_sphere=dem.material(**mat), mat_wall=dem.material(**mat),
                     mass_coef=common.MASS_COEF, wall_mass=1.0, integrator=dem.CENTERED_DIFFERENCE, history_slots=16, **kw)
    g, backend = slab.make_engine_slab(cfg, scene["walls"], scene["pos"][mine], scene["radius"][mine], mine, vel=vel[mine],
                                       omega=om[mine], capacity=int(1.6 * len(mine)) + 4096,
                                       rmax_global=float(scene["radius"].max()), meshes=scene.get("meshes"))
    drv = slab.SlabDriver(backend, rank, world, lo, hi, lag=args.lag)
    drv.rebuild()
    if args.p2p:
        drv.enable_p2p(lag=2)
    if args.mesh:
        for it in range(args.steps):
            g.set_mesh_motion(0, pos=mesh_pose(it))
            drv.step(1)
    else:
        drv.step(args.steps)
    drv.drain()
    g.sync()
    sid, p, v, w = backend.export_owned()
    ok = np.array_equal(p, rp[sid]) and np.array_equal(v, rv[sid]) and np.array_equal(w, rw[sid])
    worst = float(np.abs(p - rp[sid]).max()) if len(sid) else 0.0
    counts = torch.tensor([len(sid), int(ok), drv.stats["migrated"], drv.stats["rebuilds"]], dtype=torch.int64, device="cuda")
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    if args.mesh:
        # the wrench on the mesh is the sum over the ranks of what their own spheres exert
        w = torch.tensor(g.mesh_wrench(0)[0], device="cuda", dtype=torch.float64)
        dist.all_reduce(w)
        if rank == 0:
            err = float(np.linalg.norm(w.cpu().numpy() - wr) / np.linalg.norm(wr))
            print("mesh force: slabs", w.cpu().numpy(), "single GPU", wr, "rel err %.2e" % err)
            assert err < 1e-11
    if rank == 0:
        tot = sum(int(c[0]) for c in allc)
        print("owned per rank", [int(c[0]) for c in allc], "total", tot, "bit-identical", [int(c[1]) for c in allc],
              "migrated", [int(c[2]) for c in allc], "rebuilds", [int(c[3]) for c in allc], "worst |dp|", worst)
        assert tot == n, "spheres lost or duplicated"
        assert all(int(c[1]) == 1 for c in allc), "slab run differs from the single-GPU run"
        assert sum(int(c[2]) for c in allc) > 0, "test scene produced no migration"
        print("MGPU CHECK PASSED")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

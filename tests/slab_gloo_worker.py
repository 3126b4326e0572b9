"""World-size-2 (or more) run of the slab protocol (chrono_b200/slab.py: SlabDriver) on CPU over gloo.

The CUDA engine is replaced by `OracleBackend`: the same backend interface (extract / append / select_ghosts /
finish_rebuild / pack / unpack / step / want_rebuild / export_owned) implemented with numpy on top of the CPU oracle,
ghosts being fixed bodies whose state arrives with the halo.  Frictionless (no contact history), so the result only
depends on positions and velocities and must agree with a single-process oracle run of the whole scene to rounding
(summation order differs).  What this covers is the HOST logic of the N>1 path: slab bounds, migration at rebuilds,
ghost selection, the frozen-index per-step halo, the all-reduced rebuild decision, with and without the one-step lag.

Launched by tests/test_slab_gloo.py under torch.distributed.run; every rank writes nothing, rank 0 prints the verdict."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
import dem_common as common  # noqa: E402

REC = 11  # pos3, radius, v3, w3, id


class OracleBackend:
    device = torch.device("cpu")
    halo_doubles, ghost_doubles, migrant_doubles = 9, REC, REC

    def __init__(self, scene, ids, vel, omega, dt, skin, rmax):
        self.scene, self.dt, self.skin, self.rmax = scene, dt, skin, rmax
        self.cut = 2.0 * rmax + skin
        ids = np.asarray(ids)
        self.rec = np.concatenate([scene["pos"][ids], scene["radius"][ids, None], vel[ids], omega[ids],
                                   ids[:, None].astype(np.float64)], axis=1)
        self.ghost = np.zeros(len(ids), dtype=bool)
        self.travel = 0.0
        self.last_dx = 0.0
        self.o = None

    # ---- rebuild protocol ----
    def extract(self, lo, hi):
        own = self.rec[~self.ghost]
        x = own[:, 0]
        left, right = own[x < lo], own[x >= hi]
        self.rec = own[(x >= lo) & (x < hi)]
        self.ghost = np.zeros(len(self.rec), dtype=bool)
        return len(self.rec), torch.from_numpy(left.copy()), torch.from_numpy(right.copy())

    def append(self, buf, ghost, direction):
        a = buf.numpy().reshape(-1, REC)
        if not len(a):
            return
        self.rec = np.concatenate([self.rec, a], axis=0)
        self.ghost = np.concatenate([self.ghost, np.full(len(a), bool(ghost))])

    def select_ghosts(self, lo, hi):
        assert not self.ghost.any()
        x = self.rec[:, 0]
        self.send = [np.nonzero(x < lo + self.cut)[0], np.nonzero(x >= hi - self.cut)[0]]
        return torch.from_numpy(self.rec[self.send[0]].copy()), torch.from_numpy(self.rec[self.send[1]].copy())

    def finish_rebuild(self):
        sc = dict(self.scene)
        sc["pos"], sc["radius"], sc["n"] = self.rec[:, 0:3].copy(), self.rec[:, 3].copy(), len(self.rec)
        self.o = common.make_oracle(sc, vel=self.rec[:, 4:7].copy(), omega=self.rec[:, 7:10].copy(), dt=self.dt,
                                    force_model=po.HERTZ, tangential_mode=po.TANG_NONE, num_threads=1)
        f = self.o.first_sphere_body
        for k in np.nonzero(self.ghost)[0]:
            self.o.L.orc_set_body_fixed(self.o.h, int(f + k), 1)
        self.gidx = np.nonzero(self.ghost)[0]
        self.travel, self.last_dx = 0.0, 0.0

    # ---- per step ----
    def pack(self, d):
        r = self.rec[self.send[d]]
        return torch.from_numpy(np.concatenate([r[:, 0:3], r[:, 4:10]], axis=1).copy())

    def halo_buffer(self, d, n):
        return torch.empty((n, 9), dtype=torch.float64)

    def unpack(self, d, buf):
        a = buf.numpy()
        # ghosts were appended left first, then right
        nl = getattr(self, "_nl", None)
        g = self.gidx
        rows = g[:len(a)] if d == 0 else g[len(g) - len(a):]
        self.rec[rows, 0:3] = a[:, 0:3]
        self.rec[rows, 4:10] = a[:, 3:9]

    def step(self):
        f = self.o.first_sphere_body
        for k in self.gidx:
            self.o.set_body_state(int(f + k), pos=self.rec[k, 0:3], vel=self.rec[k, 4:7], omega=self.rec[k, 7:10])
        assert self.o.step(1) == 0
        pos, rot, vel, om = self.o.state()
        # the oracle keeps omega in the body frame; records and halo messages carry the world-frame vector
        u, w0 = rot[:, 1:], rot[:, :1]
        t = 2 * np.cross(u, om)
        om = om + w0 * t + np.cross(u, t)
        own = ~self.ghost
        new = pos[f:][own]
        self.last_dx = float(np.sqrt(((new - self.rec[own, 0:3]) ** 2).sum(axis=1)).max()) if own.any() else 0.0
        self.travel += self.last_dx
        self.rec[own, 0:3] = new
        self.rec[own, 4:7] = vel[f:][own]
        self.rec[own, 7:10] = om[f:][own]

    def want_rebuild(self, ahead=0):
        t = self.travel + ahead * self.last_dx
        return torch.tensor([0 if t < 0.499 * self.skin else 1], dtype=torch.int32)

    def flag_to_host_async(self, flag):
        v = int(flag[0])
        return lambda: v

    def export_owned(self):
        own = self.rec[~self.ghost]
        return own[:, 10].astype(np.int64), own[:, 0:3], own[:, 4:7], own[:, 7:10]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=1500)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--lag", type=int, default=0)
    args = ap.parse_args()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from chrono_b200 import scenes, slab
    n = args.spheres
    scene = scenes.settling_scene(n, sep_factor=1.99, seed=77)
    rng = np.random.default_rng(5)
    vel = rng.normal(size=(n, 3)) * 0.1
    # the outer parts drift towards the middle, across every slab face
    xq = np.quantile(scene["pos"][:, 0], [1.0 / 3, 2.0 / 3]) if world > 2 else (0.0, 0.0)
    vel[:, 0] += np.where(scene["pos"][:, 0] < xq[0], 1.5, np.where(scene["pos"][:, 0] >= xq[1], -1.5, 0.0))
    om = rng.normal(size=(n, 3)) * 2.0
    dt, rmax = 1e-4, float(scene["radius"].max())
    skin = 0.25 * rmax

    bounds = slab.slab_bounds(scene["pos"][:, 0], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = np.nonzero((scene["pos"][:, 0] >= lo) & (scene["pos"][:, 0] < hi))[0]
    b = OracleBackend(scene, mine, vel, om, dt, skin, rmax)
    drv = slab.SlabDriver(b, rank, world, lo, hi, lag=args.lag)
    drv.rebuild()
    drv.step(args.steps)
    drv.drain()
    sid, p, v, w = b.export_owned()

    ok, worst = True, 0.0
    if True:
        ref = common.make_oracle(scene, vel=vel, omega=om, dt=dt, force_model=po.HERTZ, tangential_mode=po.TANG_NONE, num_threads=1)
        assert ref.step(args.steps) == 0
        rp, _, rv, _ = ref.state()
        f = ref.first_sphere_body
        worst = float(np.abs(p - rp[f:][sid]).max()) if len(sid) else 0.0
        wv = float(np.abs(v - rv[f:][sid]).max()) if len(sid) else 0.0
        ok = worst < 1e-10 and wv < 1e-7
    t = torch.tensor([len(sid), int(ok), drv.stats["migrated"], drv.stats["rebuilds"], drv.stats["halo_bytes"]], dtype=torch.int64)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    if rank == 0:
        tot = sum(int(c[0]) for c in allc)
        print("owned", [int(c[0]) for c in allc], "ok", [int(c[1]) for c in allc], "migrated", [int(c[2]) for c in allc],
              "rebuilds", [int(c[3]) for c in allc], "halo bytes", [int(c[4]) for c in allc], "worst |dp|", worst)
        assert tot == n, "spheres lost or duplicated"
        assert all(int(c[1]) == 1 for c in allc), "slab run differs from the single-process run"
        assert sum(int(c[2]) for c in allc) > 0, "no migration happened"
        assert all(int(c[3]) >= 3 for c in allc), "too few rebuilds to mean anything"
        assert len(set(int(c[3]) for c in allc)) == 1, "ranks disagree on the rebuild steps"
        print("SLAB GLOO PASSED")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

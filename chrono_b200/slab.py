"""Slab domain decomposition of the DEM step: one process per GPU, torch.distributed for the plumbing.

No reference counterpart exists (Chrono::Dem is single-GPU; the MPI module Chrono::Distributed was removed,
CHANGELOG.md:496-501); the design follows SURVEY.md 8(e):

* 1-D slabs along x with boundaries fixed at start-up (equal sphere counts);
* every rank keeps its owned spheres plus ghosts = copies of the neighbours' spheres within 2 r_max + skin of the face;
* between neighbour-list rebuilds the ghost SET is frozen: the per-step halo is a fixed-index pack -> send/recv -> unpack
  of (pos, v, omega), 72 bytes per ghost, no sizes exchanged;
* spheres change owner only at a rebuild and take their contact history with them; no force or history exchange, because
  both partners of a contact keep (bit-identical) copies of it;
* all ranks rebuild at the same step: a 4-byte MAX all-reduce of "my Verlet skin is used up" per step.

`SlabDriver` is the protocol; it talks to a backend object (the CUDA engine through its C ABI: `EngineBackend`; the CPU
tests plug an oracle-based backend in, tests/test_slab_gloo.py) and to torch.distributed (NCCL on GPUs, gloo on CPU).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(x_all, world):
    """Slab faces with equal sphere counts: world+1 values, first -inf, last +inf.  x_all: every sphere's x."""
    xs = np.sort(np.asarray(x_all, dtype=np.float64))
    cuts = [-np.inf]
    for r in range(1, world):
        k = (len(xs) * r) // world
        cuts.append(0.5 * (xs[k - 1] + xs[k]) if 0 < k < len(xs) else xs[min(k, len(xs) - 1)])
    cuts.append(np.inf)
    return np.array(cuts)


class SlabDriver:
    def __init__(self, backend, rank, world, lo, hi, group=None, lag=0):
        """lag = 1: the "rebuild now?" answer of step i is read on the host while step i+1 is already enqueued (no host
        sync in the step loop); the backend then asks one step early (want_rebuild(ahead=1))."""
        self.b, self.rank, self.world, self.lo, self.hi, self.group = backend, rank, world, float(lo), float(hi), group
        self.lag = int(lag)
        self._pending = []
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None
        self.stats = dict(steps=0, rebuilds=0, halo_bytes=0, migrated=0, ghost_bytes=0)
        self.fresh = False  # the ghosts were just exchanged by a rebuild: no halo needed before the next step

    # ---- transport -------------------------------------------------------------------------------------------------
    def _exchange(self, to_left, to_right, width):
        """Send variable-length record arrays to both neighbours, receive theirs.  Returns (from_left, from_right)."""
        dev = self.b.device
        counts_out = torch.tensor([0 if to_left is None else to_left.shape[0], 0 if to_right is None else to_right.shape[0]],
                                  dtype=torch.int64, device=dev)
        cl, cr = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
        ops = []
        if self.left is not None:
            ops += [dist.P2POp(dist.isend, counts_out[0:1], self.left, self.group), dist.P2POp(dist.irecv, cl, self.left, self.group)]
        if self.right is not None:
            ops += [dist.P2POp(dist.isend, counts_out[1:2], self.right, self.group), dist.P2POp(dist.irecv, cr, self.right, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        nl, nr = int(cl.item()), int(cr.item())
        from_left = torch.empty((nl, width), dtype=torch.float64, device=dev)
        from_right = torch.empty((nr, width), dtype=torch.float64, device=dev)
        ops = []
        if self.left is not None:
            if to_left is not None and to_left.shape[0]:
                ops.append(dist.P2POp(dist.isend, to_left.contiguous(), self.left, self.group))
            if nl:
                ops.append(dist.P2POp(dist.irecv, from_left, self.left, self.group))
        if self.right is not None:
            if to_right is not None and to_right.shape[0]:
                ops.append(dist.P2POp(dist.isend, to_right.contiguous(), self.right, self.group))
            if nr:
                ops.append(dist.P2POp(dist.irecv, from_right, self.right, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return from_left, from_right

    # ---- rebuild: migration, new ghost set ---------------------------------------------------------------------------
    def rebuild(self):
        b = self.b
        t0 = b.clock() if hasattr(b, "clock") else None  # synchronising wall clock: rebuilds only
        if getattr(self, "p2p", False):
            # device-driven: records land in the neighbours' buffers by peer stores, one host sync for the new layout
            c = b.p2p_rebuild(self.lo, self.hi)
            self.n_recv, self.n_send = (c[1], c[2]), (c[3], c[4])
            self.stats["migrated"] += c[5] + c[6]
            self.stats["ghost_bytes"] += int((c[1] + c[2]) * b.ghost_doubles * 8)
            self.stats["rebuilds"] += 1
            self.fresh = True
            self.stats["rebuild_s"] = self.stats.get("rebuild_s", 0.0) + (b.clock() - t0)
            return
        n_keep, out_l, out_r = b.extract(self.lo, self.hi)
        if self.left is None and out_l.shape[0] or self.right is None and out_r.shape[0]:
            raise RuntimeError("a sphere left the outermost slab (bounds must be infinite there)")
        in_l, in_r = self._exchange(out_l, out_r, b.migrant_doubles)
        b.append(in_l, ghost=False, direction=0)
        b.append(in_r, ghost=False, direction=1)
        self.stats["migrated"] += int(out_l.shape[0] + out_r.shape[0])
        g_l, g_r = b.select_ghosts(self.lo, self.hi)
        if self.left is None:
            g_l = g_l[:0]
        if self.right is None:
            g_r = g_r[:0]
        gin_l, gin_r = self._exchange(g_l, g_r, b.ghost_doubles)
        self.stats["ghost_bytes"] += int((gin_l.numel() + gin_r.numel()) * 8)
        b.append(gin_l, ghost=True, direction=0)
        b.append(gin_r, ghost=True, direction=1)
        b.finish_rebuild()
        self.n_recv = (int(gin_l.shape[0]), int(gin_r.shape[0]))
        self.n_send = (int(g_l.shape[0]), int(g_r.shape[0]))
        self.stats["rebuilds"] += 1
        self.fresh = True
        if t0 is not None:
            self.stats["rebuild_s"] = self.stats.get("rebuild_s", 0.0) + (b.clock() - t0)

    # ---- per-step halo -------------------------------------------------------------------------------------------------
    def halo(self):
        b = self.b
        send = [b.pack(0) if self.left is not None else None, b.pack(1) if self.right is not None else None]
        recv = [b.halo_buffer(0, self.n_recv[0]), b.halo_buffer(1, self.n_recv[1])]
        ops = []
        for d, peer in ((0, self.left), (1, self.right)):
            if peer is None:
                continue
            if self.n_send[d]:
                ops.append(dist.P2POp(dist.isend, send[d], peer, self.group))
            if self.n_recv[d]:
                ops.append(dist.P2POp(dist.irecv, recv[d], peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for d in (0, 1):
            if self.n_recv[d]:
                b.unpack(d, recv[d])
        self.stats["halo_bytes"] += int((self.n_recv[0] + self.n_recv[1]) * b.halo_doubles * 8)

    # ---- direct P2P halo (engine backend only) -------------------------------------------------------------------------
    def enable_p2p(self, lag=2):
        """Switch the per-step path to NVLink peer stores + device-side vote (include/chrono_b200_dem.h, dem_b200_p2p_*):
        from now on one engine call per step does halo + step + vote; the host only polls the vote `lag` steps late and
        runs rebuild() when it says so.  Handles travel through torch.distributed once."""
        handle = self.b.p2p_export()
        mine = torch.from_numpy(handle.copy()).to(self.b.device)
        allh = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self.group)
        self.b.p2p_import(self.rank, self.world, torch.stack(allh).cpu().numpy(), lag)
        self.p2p, self.lag = True, int(lag)
        self.first_vote = self.b.step_count() + 1
        self.ignore_upto = 0

    def _step_p2p(self, nsteps):
        b = self.b
        for _ in range(nsteps):
            b.step()  # halo (unless the ghosts are fresh from a rebuild) + step + vote, one graph launch
            self.fresh = False
            self.stats["steps"] += 1
            self.stats["halo_bytes"] += int((self.n_recv[0] + self.n_recv[1]) * b.halo_doubles * 8)
            k = b.step_count() - self.lag
            if k >= self.first_vote and k > self.ignore_upto and b.p2p_poll_vote(k):
                self.rebuild()
                self.ignore_upto = b.step_count()  # votes cast before the rebuild are void

    def step(self, nsteps=1):
        if getattr(self, "p2p", False):
            return self._step_p2p(nsteps)
        for _ in range(nsteps):
            if not self.fresh:
                self.halo()
            self.fresh = False
            self.b.step()
            flag = self.b.want_rebuild(self.lag)  # 1-element int32 tensor on the backend's device
            if self.world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            self.stats["steps"] += 1
            if not self.lag:
                if int(flag.item()):
                    self.rebuild()
                continue
            # every rank sees the same all-reduced flags in the same order, so all take the same decision
            self._pending.append(self.b.flag_to_host_async(flag))
            if len(self._pending) > self.lag:
                if self._pending.pop(0)():
                    self._pending.clear()
                    self.rebuild()

    def import_owned(self, pos=None, vel=None, omega=None):
        """Host round trip of a slab engine: the records of the last export_owned() go back (same order), then the slab
        is rebuilt (new positions invalidate the candidate lists and the neighbours' ghost copies).  Every rank calls it."""
        self.b.import_owned(pos, vel, omega)
        self._pending.clear()
        self.rebuild()
        if getattr(self, "p2p", False):
            self.ignore_upto = self.b.step_count()

    def drain(self):
        """Act on the answers still in flight (call before reading results / at the end of a timed region)."""
        if getattr(self, "p2p", False):
            return  # votes not yet polled are looked at by the next step() call
        while self._pending:
            if self._pending.pop(0)():
                self._pending.clear()
                self.rebuild()


class EngineBackend:
    """The CUDA engine (libchrono_b200_dem.so) as a slab backend.  Communication buffers are torch tensors whose device
    pointers are handed to the C ABI; the engine runs on torch's current stream so that NCCL and the kernels are ordered."""

    def __init__(self, system, capacity, max_records=None):
        from . import dem
        self.g = system
        self.L = system.L
        self.h = system.h
        self.device = torch.device("cuda", torch.cuda.current_device())
        hb, gb, mb, cut = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_double(0)
        system._ck(self.L.dem_b200_mgpu_sizes(self.h, C.byref(hb), C.byref(gb), C.byref(mb), C.byref(cut)))
        self.halo_doubles, self.ghost_doubles, self.migrant_doubles = hb.value // 8, gb.value // 8, mb.value // 8
        self.cut = cut.value
        self.capacity = capacity
        self.max_records = max_records or max(1024, capacity // 4)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.mig = [torch.empty((max(4096, self.max_records // 8), self.migrant_doubles), **f64) for _ in range(2)]
        self.gho = [torch.empty((self.max_records, self.ghost_doubles), **f64) for _ in range(2)]
        self.hsend = [torch.empty((self.max_records, self.halo_doubles), **f64) for _ in range(2)]
        self.hrecv = [torch.empty((self.max_records, self.halo_doubles), **f64) for _ in range(2)]
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_send = [0, 0]
        self.L.dem_b200_step_count.restype = C.c_ulonglong
        self.L.dem_b200_step_count.argtypes = [C.c_void_p]
        self.L.dem_b200_p2p_poll_vote.argtypes = [C.c_void_p, C.c_ulonglong, C.POINTER(C.c_int)]

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def extract(self, lo, hi):
        nk, nl, nr = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self.g._ck(self.L.dem_b200_mgpu_extract(self.h, C.c_double(lo), C.c_double(hi), self._p(self.mig[0]), self._p(self.mig[1]),
                                                C.c_size_t(self.mig[0].shape[0]), C.byref(nk), C.byref(nl), C.byref(nr)))
        return nk.value, self.mig[0][:nl.value], self.mig[1][:nr.value]

    def append(self, buf, ghost, direction):
        n = int(buf.shape[0])
        self.g._ck(self.L.dem_b200_mgpu_append(self.h, self._p(buf) if n else None, C.c_size_t(n), int(ghost), int(direction)))

    def select_ghosts(self, lo, hi):
        nl, nr = C.c_size_t(0), C.c_size_t(0)
        self.g._ck(self.L.dem_b200_mgpu_select_ghosts(self.h, C.c_double(lo), C.c_double(hi), C.c_double(-1.0), self._p(self.gho[0]),
                                                      self._p(self.gho[1]), C.c_size_t(self.max_records), C.byref(nl), C.byref(nr)))
        self.n_send = [nl.value, nr.value]
        return self.gho[0][:nl.value], self.gho[1][:nr.value]

    def finish_rebuild(self):
        self.g._ck(self.L.dem_b200_mgpu_finish_rebuild(self.h))

    def pack(self, d):
        self.g._ck(self.L.dem_b200_mgpu_pack(self.h, d, self._p(self.hsend[d])))
        return self.hsend[d][:self.n_send[d]]

    def halo_buffer(self, d, n):
        return self.hrecv[d][:n]

    def unpack(self, d, buf):
        self.g._ck(self.L.dem_b200_mgpu_unpack(self.h, d, self._p(buf)))

    def step(self):
        self.g._ck(self.L.dem_b200_step(self.h, 1))

    def p2p_export(self):
        h = np.zeros(64, dtype=np.uint8)
        self.g._ck(self.L.dem_b200_p2p_export(self.h, C.c_size_t(self.max_records), h.ctypes.data_as(C.c_void_p)))
        return h

    def p2p_import(self, rank, world, handles, ahead):
        hs = np.ascontiguousarray(handles, dtype=np.uint8)
        self.g._ck(self.L.dem_b200_p2p_import(self.h, int(rank), int(world), hs.ctypes.data_as(C.c_void_p), int(ahead)))

    def p2p_rebuild(self, lo, hi):
        c = (C.c_size_t * 7)()
        self.g._ck(self.L.dem_b200_p2p_rebuild(self.h, C.c_double(lo), C.c_double(hi), c))
        self.n_send = [c[3], c[4]]
        return [int(x) for x in c]

    def p2p_poll_vote(self, step):
        f = C.c_int(0)
        self.g._ck(self.L.dem_b200_p2p_poll_vote(self.h, C.c_ulonglong(step), C.byref(f)))
        return f.value

    def step_count(self):
        return int(self.L.dem_b200_step_count(self.h))

    def clock(self):
        import time
        torch.cuda.synchronize()
        return time.perf_counter()

    def want_rebuild(self, ahead=0):
        self.g._ck(self.L.dem_b200_mgpu_want_rebuild_ahead(self.h, C.c_void_p(self.flag.data_ptr()), int(ahead)))
        return self.flag

    def flag_to_host_async(self, flag):
        """Copy the (all-reduced) flag to pinned host memory on the current stream; returns a callable that waits for the
        copy and gives the value."""
        if not hasattr(self, "_pin"):
            self._pin = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(8)]
            self._pin_i = 0
        h = self._pin[self._pin_i % len(self._pin)]
        self._pin_i += 1
        h.copy_(flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def wait():
            ev.synchronize()
            return int(h[0])
        return wait

    def import_owned(self, pos, vel, omega):
        a = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None for x in (pos, vel, omega)]
        n = next(len(x) for x in a if x is not None)
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double)) if x is not None else None
        self.g._ck(self.L.dem_b200_import_owned(self.h, C.c_size_t(n), dp(a[0]), dp(a[1]), dp(a[2])))

    def export_owned(self):
        cap = self.capacity
        sid = np.empty(cap, dtype=np.uint32)
        pos, vel, om = np.empty((cap, 3)), np.empty((cap, 3)), np.empty((cap, 3))
        n = C.c_size_t(0)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self.g._ck(self.L.dem_b200_export_owned(self.h, sid.ctypes.data_as(C.POINTER(C.c_uint32)), dp(pos), dp(vel), dp(om),
                                                C.c_size_t(cap), C.byref(n)))
        k = n.value
        return sid[:k], pos[:k], vel[:k], om[:k]


def slab_parity_check(rank, world, local, n=200000, steps=300, p2p=True, seed=77):
    """A slab run of `n` spheres over `world` ranks against the same job on ONE GPU (every rank runs the whole scene on its
    own device with the plain engine): positions, velocities and angular velocities of the owned spheres must be bit-identical
    and the number of force-carrying contacts equal.  The spheres drift across the slab faces, so migration (with contact
    history) and several list rebuilds are part of it.  Collective: every rank calls it; returns the same dict everywhere."""
    from . import dem, scenes
    scene = scenes.settling_scene(n, sep_factor=1.99, seed=seed)
    rng = np.random.default_rng(5)
    vel = rng.normal(size=(n, 3)) * 0.1
    vel[:, 0] += np.where(scene["pos"][:, 0] < 0, 0.6, -0.6)  # both halves drift towards (and across) the middle
    om = rng.normal(size=(n, 3)) * 2.0
    kw = dict(dt=1e-4, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP)
    ref = scenes.make_gpu(scene, vel=vel, omega=om, device=local, **kw)
    ref.step(steps)
    rp, rv, rw = ref.state()
    contacts_ref = ref.reduce(dem.RED_NUM_CONTACTS)
    ref.close()
    x = scene["pos"][:, 0]
    bounds = slab_bounds(x, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = np.nonzero((x >= lo) & (x < hi))[0]
    mat = scenes.settling_material()
    cfg = dem.config(device=local, bins=scene["bins"], mat_sphere=dem.material(**mat), mat_wall=dem.material(**mat),
                     mass_coef=scenes.MASS_COEF, wall_mass=1.0, integrator=dem.CENTERED_DIFFERENCE, history_slots=16, **kw)
    g, backend = make_engine_slab(cfg, scene["walls"], scene["pos"][mine], scene["radius"][mine], mine, vel=vel[mine],
                                  omega=om[mine], capacity=int(1.6 * len(mine)) + 4096, rmax_global=float(scene["radius"].max()))
    drv = SlabDriver(backend, rank, world, lo, hi, lag=0)
    drv.rebuild()
    if p2p:
        drv.enable_p2p(lag=2)
    drv.step(steps)
    drv.drain()
    g.sync()
    sid, p, v, w = backend.export_owned()
    ok = bool(np.array_equal(p, rp[sid]) and np.array_equal(v, rv[sid]) and np.array_equal(w, rw[sid]))
    worst = float(np.abs(p - rp[sid]).max()) if len(sid) else 0.0
    t = torch.tensor([len(sid), int(ok), drv.stats["migrated"], g.reduce(dem.RED_NUM_CONTACTS), worst], dtype=torch.float64,
                     device=backend.device)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    g.close()
    tot = int(sum(float(a[0]) for a in allt))
    contacts = int(sum(float(a[3]) for a in allt))
    return dict(bit_identical=bool(all(int(a[1]) == 1 for a in allt)) and tot == n, worst_dp=max(float(a[4]) for a in allt),
                spheres=n, timesteps=steps, owned_total=tot, migrated=int(sum(float(a[2]) for a in allt)),
                rebuilds=int(drv.stats["rebuilds"]), contacts_slabs=contacts, contacts_single_gpu=int(contacts_ref),
                contacts_equal=contacts == int(contacts_ref), transport="p2p" if p2p else "nccl")


def make_engine_slab(cfg, scene_walls, pos, radius, ids, vel=None, omega=None, capacity=None, rmax_global=None, add_walls=None,
                     meshes=None):
    """Create a slab-mode engine for the given owned spheres (global ids `ids`) on the current CUDA device/stream."""
    from . import dem
    g = dem.DemSystem(cfg)
    L = g.L
    g._ck(L.dem_b200_set_stream(g.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for p, h in scene_walls:
        g.add_box_wall(p, h)
    if add_walls:
        add_walls(g)
    for M in (meshes or []):  # meshes (and walls) are replicated on every rank; each rank keeps the wrench of ITS spheres
        m = g.add_mesh(M["tri"], M.get("mass", 1.0))
        g.set_mesh_motion(m, M.get("pos"), M.get("rot"), M.get("vel"), M.get("omega"))
    n = len(ids)
    capacity = capacity or int(1.5 * n + 4096)
    g.set_spheres(pos, radius, vel=vel, omega=omega)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    g._ck(L.dem_b200_set_sphere_ids(g.h, ids.ctypes.data_as(C.POINTER(C.c_uint32))))
    g._ck(L.dem_b200_mgpu_enable(g.h, C.c_size_t(capacity), C.c_double(rmax_global if rmax_global is not None else float(np.max(radius)))))
    g.initialize()
    return g, EngineBackend(g, capacity)

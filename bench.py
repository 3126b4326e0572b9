#!/usr/bin/env python
"""Benchmark of the SMC granular DEM hot path (sphere-steps/s), bench contract of the graft driver.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spheres M] [--substeps S]

Workload (BASELINE.json configs[1]): M = 1,000,000 monodisperse spheres (R = 0.02 m, rho = 2000, Y = 2e6, mu = 0.4,
COR = 0.4), Hertz-Mindlin with MultiStep tangential history, five-wall box, h = 1e-4 s.  The packing is synthetic:
a jittered HCP lattice at spacing 2R, i.e. about half of the 12 lattice neighbours overlap -> c_bar ~ 6 contacts per
sphere from the first step, identical for the GPU arm and the CPU reference arm (no settling phase needed).

One bench "step" = one AdvanceSimulation-style call of S DEM time steps (default 100, a typical output-frame
cadence).  value = M*S*K / device time with the state resident in HBM; e2e = same through the host-buffer C-ABI
call dem_b200_advance_host (H2D of pos/vel/omega + S steps + D2H, every bench step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RADIUS = 0.02
DT = 1e-4
COND_GRAPH = os.environ.get("DEMB200_COND_GRAPH", "")[:1] == "1"  # opt-in: see build_graph() in csrc/dem_engine.cu
KERNELS_PER_TIMESTEP = 9  # launches of our own kernels per DEM time step (7 of them return at once unless the step rebuilds)


POLY = None      # --polydisperse lo,hi : radii U(lo, hi) * R (BASELINE configs[2]-style physics); not the default workload
MU_ROLL = 0.0    # --mu-roll
USER_COEFF = False  # --user-coeff


def build_scene(n):
    from chrono_b200 import scenes
    if POLY:
        return scenes.settling_scene(n, radius=RADIUS, sep_factor=1.75, jitter=0.005, seed=12345 + 2, polydisperse=POLY)
    return scenes.settling_scene(n, radius=RADIUS, sep_factor=2.0, jitter=0.005, seed=12345 + 1)


def build_scene_slabs(n_per_gpu, world):
    """Weak scaling: the single-GPU box stretched `world` times along x (same depth and width of the bed), world * n
    spheres, cut into `world` slabs of equal sphere count."""
    from chrono_b200 import scenes
    one = build_scene(n_per_gpu)
    Lx, Ly = one["box_size"][0], one["box_size"][1]
    return scenes.settling_scene(n_per_gpu * world, radius=RADIUS, sep_factor=2.0, jitter=0.005, seed=12345 + 1,
                                 box_xy=(Lx * world, Ly))


def measured_traffic(n, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload), else None."""
    p = os.path.join(ROOT, "profiles", "force_traffic.json")
    if n != 1000000 or not kernel.startswith("k_force") or not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    return float(t["dram_bytes_read"] + t["dram_bytes_write"])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: do not let that shrink the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_run(scene, timesteps, threads=0):
    """Times the CPU oracle (restatement of Chrono::Multicore SMC) on the same packing for `timesteps` steps."""
    import dem_common as common
    from oracle import pyoracle as po
    threads = threads or host_threads()
    o = common.make_oracle(scene, dt=DT, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP, num_threads=threads)
    o.step(1)  # untimed: first-touch of all arrays
    o.reset_timers()
    t0 = time.perf_counter()
    rc = o.step(timesteps)
    t1 = time.perf_counter()
    assert rc == 0
    return scene["n"] * timesteps / (t1 - t0), threads, o.timers()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port of ChSystemMulticoreSMC; the real
    class cannot be built here because Chrono core needs Eigen3), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import dem_common as common
    from oracle import pyoracle as po
    n = args.spheres
    scene = build_scene(n)
    sample_steps = args.ref_substeps
    cores = host_threads()
    o = common.make_oracle(scene, dt=DT, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP, num_threads=cores)
    for _ in range(args.warmup):
        o.step(sample_steps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(sample_steps)
    t1 = time.perf_counter()
    val = n * sample_steps * args.steps / (t1 - t0)
    line = {
        "impl": "reference", "metric": "sphere-steps/sec", "value": val, "unit": "sphere-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (t1 - t0) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, sample_steps, args.gpus),
        "cpu_baseline": {"value": val, "unit": "sphere-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d spheres x %d time steps per bench step (same packing as the GPU arm)" % (n, sample_steps)},
        "e2e": {"value": val, "unit": "sphere-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n, substeps, gpus):
    return {"workload": ("BASELINE configs[1]: %d spheres Hertz-Mindlin MultiStep history in a 5-wall box, jittered HCP "
                         "packing at 2R spacing (c_bar~6), R=0.02 rho=2000 Y=2e6 mu=0.4 cr=0.4 h=1e-4" % n) +
                        ((" -- VARIANT: radii U(%g,%g) R, mu_roll %g" % (POLY[0], POLY[1], MU_ROLL)) if POLY else
                         ((" -- VARIANT: mu_roll %g" % MU_ROLL) if MU_ROLL else "")) + (" -- VARIANT: user coefficients" if USER_COEFF else ""),
            "spheres_per_gpu": n, "timesteps_per_step": substeps,
            "l2_policy": "working set (>=300 B/sphere x %d spheres) exceeds the 126 MB L2; no explicit flush" % n,
            "step_graph": ("conditional rebuild node (DEMB200_COND_GRAPH=1; its kernels are invisible to ncu)"
                           if gpus == 1 and COND_GRAPH else "flat capture, every launch profilable"),
            "parallelism": "1 process per GPU" if gpus == 1 else
            "slab domain decomposition along x, %d ranks x %d spheres (box %d x as long), ghost halo over NVLink" % (gpus, n, gpus)}


def bind_to_gpu_numa_node(local):
    """One process per GPU: keep the host thread (and the threads NCCL / CUDA spawn) on the cores next to that GPU
    (`nvidia-smi topo -m`, column "CPU Affinity")."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        lines = [l for l in out.splitlines() if l.strip()]
        hdr = next(l for l in lines if "CPU Affinity" in l)
        cols = [c.strip() for c in hdr.split("\t") if c.strip()]
        ci = cols.index("CPU Affinity") + 1  # rows carry the row label in front
        row = next(l for l in lines if l.startswith("GPU%d" % local))
        cells = [c.strip() for c in row.split("\t") if c.strip()]
        spec = cells[ci]
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


def run_slabs(args):
    """N > 1: slab domain decomposition (chrono_b200/slab.py).  One process per GPU, every rank owns n spheres of a box
    N times as long; per step: ghost halo (pos, v, omega of the spheres within 2 r_max + skin of a slab face) over NCCL
    stores, the same step graph as on one GPU, a device-side vote on "rebuild now?"; spheres migrate at rebuilds (NCCL)."""
    import torch
    import torch.distributed as dist
    from chrono_b200 import dem, slab
    import dem_common as common

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_to_gpu_numa_node(local)
    # NCCL writes its version banner (NCCL_DEBUG >= VERSION) to stdout: send its log to stderr, rank 0 prints ONE JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, S = args.spheres, args.substeps
    scene = build_scene_slabs(n, world)
    x = scene["pos"][:, 0]
    bounds = slab.slab_bounds(x, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = np.nonzero((x >= lo) & (x < hi))[0]
    mat = common.settling_material()
    cfg = dem.config(device=local, dt=DT, bins=scene["bins"], mat_sphere=dem.material(**mat), mat_wall=dem.material(**mat),
                     mass_coef=common.MASS_COEF, wall_mass=1.0, integrator=dem.CENTERED_DIFFERENCE, history_slots=16,
                     force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP, verlet_skin=args.skin * RADIUS if args.skin > 0 else -1.0)
    g, backend = slab.make_engine_slab(cfg, scene["walls"], scene["pos"][mine], scene["radius"][mine], mine,
                                       capacity=int(1.15 * len(mine)) + 65536, rmax_global=float(scene["radius"].max()))
    drv = slab.SlabDriver(backend, rank, world, lo, hi, lag=1)
    drv.rebuild()
    if not args.nccl_halo:
        drv.enable_p2p(lag=args.slab_lag)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    for _ in range(args.warmup):
        drv.step(S)
    drv.drain()
    g.sync()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:  # one nvidia-smi poller per job: the queries take driver locks that every rank's launches contend for
        sampler.start()
    st0 = dict(drv.stats)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        drv.step(S)
    drv.drain()
    ev1.record()
    t_enqueue = time.perf_counter() - t_wall0  # host time to issue the work (close to t_wall = the host is the bottleneck)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    rebuild_s_timed = drv.stats.get("rebuild_s", 0.0) - st0.get("rebuild_s", 0.0)
    g.sync()  # device error flags (skin exceeded, overflow, NaN) fail loudly here
    t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    nsteps = args.steps * S
    value = world * n * nsteps / (ms_total * 1e-3)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    # ---- e2e: every bench step the owned state comes from / goes back to pinned host memory
    sid, p0, v0, w0 = backend.export_owned()
    n_own = len(sid)
    bytes_io = int(3 * p0.nbytes)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        drv.step(S)
        drv.drain()
        sid, p0, v0, w0 = backend.export_owned()  # D2H of the owned spheres (ids + pos + vel + omega)
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * S * e2e_steps / float(te.item())

    halo = torch.tensor([drv.stats["halo_bytes"] - st0["halo_bytes"], drv.stats["migrated"] - st0["migrated"],
                         drv.stats["rebuilds"] - st0["rebuilds"], n_own, rebuild_s_timed], device="cuda", dtype=torch.float64)
    allh = [torch.zeros_like(halo) for _ in range(world)]
    dist.all_gather(allh, halo)
    rows = g.reduce(dem.RED_NUM_CONTACTS)
    cb = torch.tensor([rows], device="cuda", dtype=torch.float64)
    dist.all_reduce(cb)
    cbar = float(cb.item()) / (world * n)
    hbm, hbm_src = peaks()
    B_step = 176.0 + 32.0 * cbar
    if rank == 0:
        halo_rank_max = max(float(h[0]) for h in allh)
        line = {
            "metric": "sphere-steps/sec", "value": value, "unit": "sphere-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, S, world),
            "contacts_per_sphere": cbar,
            "roofline": {"bound": "hbm", "achieved": B_step * (value / world) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": B_step * (value / world) / 1e9 / hbm, "traffic": None, "kernel": "whole step, per GPU",
                         "peak_source": hbm_src, "algorithmic_bytes_per_sphere": B_step},
            "halo": {"bytes_received_per_rank_per_timestep": [float(h[0]) / nsteps for h in allh],
                     "nvlink_GBps_busiest_rank": halo_rank_max / (ms_total * 1e-3) / 1e9, "nvlink_peak_GBps_per_direction": 900.0,
                     "migrated_spheres": [int(h[1]) for h in allh], "rebuilds": int(allh[0][2]),
                     "owned_spheres": [int(h[3]) for h in allh],
                     "rebuild_ms_total_in_timed_region": [1e3 * float(h[4]) for h in allh],
                     "ms_per_timestep_outside_rebuilds": (ms_total - 1e3 * max(float(h[4]) for h in allh)) / nsteps, "transport": "NCCL send/recv + 4-byte all-reduce per time step" if args.nccl_halo else
                     "NVLink peer stores from the pack kernel (CUDA IPC), device-side flags and vote; no collective per time step"},
            "cpu_baseline": None,
            "e2e": {"value": e2e_value, "unit": "sphere-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": bytes_io,
                    "steps": e2e_steps, "note": "slab mode: owned state read back to the host every bench step; the state stays "
                                                "resident on the GPUs between steps (uploading it would re-partition the domain)"},
            "gpu_launches": int(nsteps * (KERNELS_PER_TIMESTEP + 5)),
            "clocks": sampler.result(), "wall_s_timed_region": t_wall, "host_enqueue_s": t_enqueue, "cpu_affinity_rank0": affinity,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from chrono_b200 import dem
    import dem_common as common

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not args.replicas:
        return run_slabs(args)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    n = args.spheres
    S = args.substeps
    scene = build_scene(n)
    extra = dict(mat=common.settling_material(mu_roll=MU_ROLL), history_slots=24) if (POLY or MU_ROLL) else {}
    if USER_COEFF:  # the model every in-tree Chrono::Dem caller uses: explicit kn / kt / gn / gt (SetKn_SPH2SPH ...)
        extra["use_mat_props"] = False
        m = common.settling_material(mu_roll=MU_ROLL)
        m.update(kn=2e7, kt=2e7, gn=40.0, gt=20.0)
        extra["mat"] = m
    g = common.make_gpu(scene, dt=DT, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP, **extra)
    g.L.dem_b200_step  # the CUDA extension is loaded; there is no other path

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up
    for _ in range(args.warmup):
        g.step(S, sync=False)
    g.sync()
    barrier()
    # ---- timed region: K bench steps, device time from CUDA events on the engine's own stream
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ms_total += g.step_timed(S)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * n * S * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host buffers out, every bench step
    pos, vel, om = g.state()
    hp = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (pos, vel, om)]
    ho = [torch.empty(a.shape, dtype=torch.float64).pin_memory().numpy() for a in (pos, vel, om)]
    g.advance_host(hp[0], hp[1], hp[2], S, ho[0], ho[1], ho[2])  # warm
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.advance_host(hp[0], hp[1], hp[2], S, ho[0], ho[1], ho[2])
        hp, ho = ho, hp  # the next call starts from this call's result: the caller hands the output buffers back in
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = world * n * S * e2e_steps / t_e2e
    sampler.stop_flag = True
    sampler.join(timeout=2)
    bytes_io = int(3 * pos.nbytes)

    # ---- per-kernel split of one time step (events around every launch) and contact statistics
    prof = g.step_profile(20)
    prof = {k: v / 20 for k, v in prof.items()}
    # force-carrying contacts per sphere (sphere-sphere counted on both partners, sphere-wall once)
    rows = g.reduce(dem.RED_NUM_CONTACTS)
    cbar = rows / n
    stats = g.stats()
    hbm, hbm_src = peaks()
    dom = max(prof, key=prof.get)
    # algorithmic bytes of the dominant kernel (narrowphase + force + integrate): state R+W 144, radius/id 16,
    # history 32*c_bar (SURVEY 8d minus the 16 B/sphere that belong to the sort kernels)
    B_kernel = 160.0 + 32.0 * (cbar or 6.0)
    B_step = 176.0 + 32.0 * (cbar or 6.0)
    achieved = B_kernel * n / (prof[dom] * 1e-3) / 1e9
    step_ms = sum(prof.values())

    if rank == 0:
        cpu_val, cpu_cores, cpu_t = cpu_oracle_run(scene, args.cpu_steps) if args.cpu_steps > 0 else (None, 0, {})
        # per-core figure: the same port on ONE thread, one time step of the same packing
        cpu_1t = cpu_oracle_run(scene, 1, threads=1)[0] if args.cpu_steps > 0 else None
        line = {
            "metric": "sphere-steps/sec", "value": value, "unit": "sphere-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, S, world),
            "contacts_per_sphere": cbar, "history_records": rows,
            "neighbor_list_rebuilds": stats["rebuilds"], "timesteps_total": stats["steps"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": measured_traffic(n, dom), "traffic_source": "profiles/force_traffic.json (ncu, bytes per launch)",
                         "kernel": dom, "kernel_ms": prof[dom], "peak_source": hbm_src,
                         "algorithmic_bytes_per_sphere": B_kernel,
                         "whole_step_frac": B_step * (value / world) / 1e9 / hbm},
            "kernel_ms_per_timestep": prof, "kernel_share": {k: v / step_ms for k, v in prof.items()},
            "cpu_baseline": {"value": cpu_val, "unit": "sphere-steps/s", "cores": cpu_cores, "kind": "port",
                             "sample": "%d spheres x %d time steps of the same packing" % (n, args.cpu_steps),
                             "phase_seconds": cpu_t, "value_one_thread": cpu_1t},
            "e2e": {"value": e2e_value, "unit": "sphere-steps/s", "h2d_bytes_per_step": bytes_io,
                    "d2h_bytes_per_step": bytes_io, "steps": e2e_steps},
            # conditional step graph: only k_step_begin + k_force_integrate launch in a step that does not rebuild
            # (the seven rebuild launches of the rebuilding steps are left out: a lower bound)
            "gpu_launches": int(args.steps * S * (2 if COND_GRAPH and world == 1 else KERNELS_PER_TIMESTEP)),
            "clocks": sampler.result(), "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spheres", type=int, default=1000000, help="spheres per GPU")
    ap.add_argument("--substeps", type=int, default=100, help="DEM time steps per bench step")
    ap.add_argument("--ref-substeps", type=int, default=2, help="time steps per bench step of the reference arm")
    ap.add_argument("--cpu-steps", type=int, default=4, help="time steps of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: NCCL send/recv halo + all-reduce vote instead of P2P stores")
    ap.add_argument("--slab-lag", type=int, default=3, help="N > 1: steps between casting the rebuild vote and acting on it")
    ap.add_argument("--skin", type=float, default=0.0, help="N > 1: Verlet skin in sphere radii (0 = engine default 0.25)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas instead of slab decomposition")
    ap.add_argument("--polydisperse", default="", help="lo,hi: radii U(lo,hi)*R instead of monodisperse (configs[2]-style; N = 1 only)")
    ap.add_argument("--user-coeff", action="store_true", help="explicit kn/kt/gn/gt instead of material properties (N = 1 only)")
    ap.add_argument("--mu-roll", type=float, default=0.0, help="rolling friction coefficient (configs[2]-style; N = 1 only)")
    args = ap.parse_args()
    global POLY, MU_ROLL, USER_COEFF
    USER_COEFF = args.user_coeff
    if args.polydisperse:
        POLY = tuple(float(x) for x in args.polydisperse.split(","))
    MU_ROLL = args.mu_roll
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

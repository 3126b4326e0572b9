#!/usr/bin/env python
"""Benchmark of the SMC granular DEM hot path (sphere-steps/s), bench contract of the graft driver.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4] [--spheres M] [--substeps S]

Workloads = BASELINE.json `configs` (R = 0.02 m, rho = 2000, Y = 2e6, mu = 0.4, COR = 0.4, h = 1e-4 s everywhere):
  --config 1  (default at N = 1) 1 M monodisperse spheres, Hertz-Mindlin with MultiStep tangential history, five-wall box
  --config 4  (default at N > 1) box fill, slab decomposition along x, weak scaling 32 M spheres per GPU (64 M @ 2 ... 256 M @ 8)
              plus a strong-scaling record (64 M spheres in total) in the same JSON line
  --config 0  10 k spheres x 1000 time steps, the GPU and the CPU port both full length
  --config 2  4 M polydisperse spheres U(0.8, 1.2) R with rolling friction 0.05 (1 GPU, or slabs at N = 2)
  --config 3  8 M spheres in a rotating drum of triangle-mesh walls, ApplyMeshMotion every time step (1 GPU, or slabs at N = 4)
The packing is synthetic (jittered HCP lattice at spacing 2R); before anything is timed the bed is SETTLED for `--settle`
time steps (reference procedure: btest_MCORE_settling.cpp:158-166 hot-starts 500 steps, then times), so the timed region sees
a stationary bed; the 1-GPU default line also carries a `flowing` record (the settled bed sheared at 1 m/s).

One bench "step" = one AdvanceSimulation-style call of S DEM time steps (default 100, a typical output-frame cadence).
value = spheres * S * K / device time with the state resident in HBM; e2e = the same through host buffers: H2D of
pos / vel / omega + S steps + D2H, every bench step (dem_b200_advance_host on one GPU; import_owned / export_owned on slabs).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RADIUS = 0.02
DT = 1e-4
COND_GRAPH = os.environ.get("DEMB200_COND_GRAPH", "1")[:1] != "0"  # default: see build_graph() in csrc/dem_engine.cu
KERNELS_PER_TIMESTEP = 9   # launches of our own kernels per DEM time step (7 of them return at once unless the step rebuilds)
KERNELS_PER_TIMESTEP_SLAB = 15  # + second pass of the force kernel (spheres that touch a ghost), 2 x halo unpack, 2 x halo pack, vote
CONFIG_NAMES = {
    0: "BASELINE configs[0]: 10k monodisperse spheres Hertz SMC settling in a box, 1000 time steps, GPU vs the CPU port full length",
    1: "BASELINE configs[1]: 1M spheres Hertz-Mindlin with MultiStep tangential history settling in a 5-wall box",
    2: "BASELINE configs[2]: 4M polydisperse spheres U(0.8,1.2)R with rolling friction 0.05 (angle-of-repose physics)",
    3: "BASELINE configs[3]: 8M spheres in a rotating drum with triangle-mesh walls (sphere-triangle contact), ApplyMeshMotion every time step",
    4: "BASELINE configs[4]: box fill, slab domain decomposition along x with NVLink halo exchange",
}
DEFAULT_SPHERES_TOTAL = {0: 10000, 1: 1000000, 2: 4000000, 3: 8000000}
WEAK_PER_GPU = 32000000
STRONG_TOTAL = 64000000
BED_LAYERS = 44  # HCP layers of the 1 M-sphere bed of configs[1] (0.26 x as deep as wide); larger beds are wider, not deeper


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(n, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the same workload, else None."""
    p = os.path.join(ROOT, "profiles", "force_traffic.json")
    if not (990000 <= n <= 1010000) or not kernel.startswith("k_force") or not os.path.exists(p):  # the capture is of configs[1]
        return None, None
    with open(p) as f:
        t = json.load(f)
    return float(t["dram_bytes_read"] + t["dram_bytes_write"]), t.get("capture", "profiles/force_traffic.json")


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


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def physics(cfg_id):
    """Engine keyword arguments (material, slots) of a config."""
    from chrono_b200 import dem, scenes
    kw = dict(dt=DT, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP)
    if cfg_id == 2:
        kw.update(mat=scenes.settling_material(mu_roll=0.05), history_slots=24)
    if cfg_id == 3:
        kw.update(history_slots=24, neighbor_slots=48)
    return kw


def build_scene(cfg_id, n):
    """Single-GPU scene of a config (the whole packing)."""
    from chrono_b200 import scenes
    if cfg_id == 2:
        return scenes.settling_scene(n, radius=RADIUS, sep_factor=1.75, jitter=0.005, seed=12345 + 2, polydisperse=(0.8, 1.2))
    if cfg_id == 3:
        return scenes.drum_scene(n, radius=RADIUS)
    if cfg_id == 0:
        return scenes.settling_scene(n, radius=RADIUS, sep_factor=2.0, jitter=0.005, seed=12345 + 1)
    # configs[1] / [4]: HCP bed of BED_LAYERS layers in static equilibrium under its own weight (it starts at rest)
    sc = scenes.slab_lattice_scene(n, world=1, rank=0, radius=RADIUS, layers=BED_LAYERS if n >= 200000 else None, precompress=True)
    sc["n"] = sc["n_total"]
    return sc


PACKING = {
    0: "jittered HCP lattice at 2R spacing in a five-wall box",
    1: "jittered HCP bed in a five-wall box, 2R in-plane spacing, layer heights in static equilibrium under gravity (the bed starts at rest)",
    2: "jittered lattice of polydisperse spheres in a five-wall box, relaxing under gravity (a slowly creeping bed: lists are rebuilt every ~25 steps)",
    3: "HCP bed in the lower third of a closed drum (cylinder wall = triangle mesh, end caps = box walls) turning at 1 m/s wall speed; the bed rides up the wall",
    4: "jittered HCP bed in a five-wall box, 2R in-plane spacing, layer heights in static equilibrium under gravity (the bed starts at rest), 44 layers deep at every size",
}


def workload_config(cfg_id, n_per_gpu, n_total, substeps, gpus, settle, extra=None):
    d = {"workload": "%s -- %d spheres%s; %s; %d untimed time steps before the timed region; "
                     "R=0.02 rho=2000 Y=2e6 mu=0.4 cr=0.4 h=1e-4" % (CONFIG_NAMES[cfg_id], n_total,
                                                                     (" (%d per GPU)" % n_per_gpu) if gpus > 1 else "", PACKING[cfg_id], settle),
         "config_id": cfg_id, "spheres_total": n_total, "spheres_per_gpu": n_per_gpu, "timesteps_per_step": substeps,
         "settle_timesteps": settle,
         "l2_policy": "working set (>= 900 B/sphere-step of DRAM traffic x %d spheres) exceeds the 126 MB L2; no explicit flush" % n_per_gpu,
         "step_graph": ("conditional rebuild node: a step that does not rebuild launches 2 kernels (DEMB200_COND_GRAPH=0 -> flat capture "
                        "of all 9, the form ncu can profile)" if gpus == 1 and COND_GRAPH else "flat capture, every launch profilable"),
         "parallelism": "1 process per GPU" if gpus == 1 else
         "slab domain decomposition along x, %d ranks, ghost halo by NVLink peer stores, spheres migrate at list rebuilds" % gpus}
    if extra:
        d.update(extra)
    return d


class DrumMotion:
    """ApplyMeshMotion of BASELINE configs[3]: the drum turns about its axis (x) so that its wall moves at `wall_speed`."""

    def __init__(self, scene, wall_speed=1.0):
        self.omega = wall_speed / scene["drum_radius"]
        self.t = 0.0

    def apply(self, g):
        from chrono_b200 import scenes
        self.t += DT
        q = scenes.quat_from_axis_angle((1.0, 0.0, 0.0), self.omega * self.t)
        g.set_mesh_motion(0, pos=np.zeros(3), rot=q, lin_vel=np.zeros(3), ang_vel=np.array([self.omega, 0.0, 0.0]))


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port; test infrastructure used as the reported baseline only)
# ---------------------------------------------------------------------------------------------------------------------
def make_cpu(scene, threads, pos=None, vel=None, omega=None, **kw):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dem_common as common
    from oracle import pyoracle as po
    sc = dict(scene)
    if pos is not None:
        sc["pos"] = pos
    model = dict(force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
    return common.make_oracle(sc, dt=DT, vel=vel, omega=omega, mat=kw.get("mat"), num_threads=threads, **model)


def cpu_oracle_run(scene, timesteps, threads=0, state=None, **kw):
    """Times the CPU oracle (restatement of Chrono::Multicore SMC) for `timesteps` steps, from `state` (pos, vel, omega:
    the GPU's settled bed) when given, else from the initial packing."""
    threads = threads or host_threads()
    o = make_cpu(scene, threads, *(state or (None, None, None)), **kw)
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
    cfg_id = args.config
    n = args.spheres or (DEFAULT_SPHERES_TOTAL.get(cfg_id, 1000000))
    n = min(n, 1000000)  # bounded sample: at most a 1 M-sphere volume of the same packing
    scene = build_scene(1 if cfg_id in (3, 4) else cfg_id, n)
    sample_steps = args.ref_substeps
    cores = host_threads()
    o = make_cpu(scene, cores, **physics(cfg_id))
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
        "config": workload_config(cfg_id, n, n, sample_steps, args.gpus, 0,
                                  {"note": "CPU arm: a %d-sphere volume of the GPU arm's packing from its initial state (configs[1] / [4]: the bed "
                                           "is generated in static equilibrium, so the contacts per sphere are those of the GPU's timed region; "
                                           "configs[2]: before the bed has relaxed, i.e. fewer contacts than the GPU arm sees)" % n}),
        "cpu_baseline": {"value": val, "unit": "sphere-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d spheres x %d time steps per bench step (same packing as the GPU arm, from its initial state)" % (n, sample_steps)},
        "e2e": {"value": val, "unit": "sphere-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_incumbent(n, timeout=240):
    """The unmodified Chrono::Dem CUDA (baseline/_ref/incumbent_dem, built by baseline/Makefile from the reference's own
    sources for sm_100a) on the same workload shape, same box.  Returns its JSON record or a reason."""
    exe = os.path.join(ROOT, "baseline", "_ref", "incumbent_dem")
    if not os.path.exists(exe):
        return {"unavailable": "baseline/_ref/incumbent_dem not built (needs /root/reference at build time)"}
    try:
        r = subprocess.run([exe, "--spheres", str(n), "--warmup", "200", "--steps", "300"], capture_output=True, text=True,
                           timeout=timeout)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": "incumbent_dem failed (rc %d): %s" % (r.returncode, (r.stderr or r.stdout)[-300:])}
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return {"unavailable": "incumbent_dem did not finish within %d s" % timeout}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": "incumbent_dem: %r" % (e,)}


# ---------------------------------------------------------------------------------------------------------------------
# one GPU
# ---------------------------------------------------------------------------------------------------------------------
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


def timed_single(g, S, K, motion=None):
    """K bench steps of S time steps; device time from CUDA events on the engine's own stream (ms)."""
    if motion is None:
        return sum(g.step_timed(S) for _ in range(K))
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K * S):
        motion.apply(g)
        g.step(1, sync=False)
    e1.record()
    e1.synchronize()
    g.sync()
    return e0.elapsed_time(e1)


def advance(g, S, motion=None):
    if motion is None:
        g.step(S, sync=False)
    else:
        for _ in range(S):
            motion.apply(g)
            g.step(1, sync=False)


def measure_single(cfg_id, n, args, local=0, with_profile=True, tag=None):
    """Build, settle, warm up, time K bench steps, e2e, kernel split.  Returns (record, engine, scene, motion)."""
    import torch
    from chrono_b200 import dem, scenes
    S = args.substeps
    scene = build_scene(cfg_id, n)
    n = scene["n"]
    kw = physics(cfg_id)
    if args.skin > 0:
        kw["verlet_skin"] = args.skin * RADIUS  # fixed (non-adaptive) Verlet skin, as slab mode uses
    motion = None
    if cfg_id == 3:  # per-step host calls (ApplyMeshMotion): run the engine on torch's stream so that torch events time it
        kw["stream"] = torch.cuda.current_stream().cuda_stream
    g = scenes.make_gpu(scene, device=local, **kw)
    if cfg_id == 3:
        motion = DrumMotion(scene)
    g.L.dem_b200_step  # the CUDA extension is loaded; there is no other path
    ke0 = g.reduce(dem.RED_KE) / n
    t0 = time.perf_counter()
    done = 0
    while done < args.settle:
        advance(g, min(500, args.settle - done), motion)
        done += min(500, args.settle - done)
    g.sync()
    settle_s = time.perf_counter() - t0
    ke1, vmax = g.reduce(dem.RED_KE) / n, g.reduce(dem.RED_MAX_SPEED)
    for _ in range(args.warmup):
        advance(g, S, motion)
    g.sync()
    torch.cuda.synchronize()
    st0 = g.stats()
    sampler = ClockSampler(local)
    sampler.start()
    t_wall0 = time.perf_counter()
    ms_total = timed_single(g, S, args.steps, motion)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    st1 = g.stats()
    value = n * S * args.steps / (ms_total * 1e-3)
    # ---- e2e: host buffers in, host buffers out, every bench step
    pos, vel, om = g.state()
    hp = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (pos, vel, om)]
    ho = [torch.empty(a.shape, dtype=torch.float64).pin_memory().numpy() for a in (pos, vel, om)]

    def e2e_step():
        nonlocal hp, ho
        if motion is None:
            g.advance_host(hp[0], hp[1], hp[2], S, ho[0], ho[1], ho[2])
        else:
            g.set_state(hp[0], hp[1], hp[2])
            advance(g, S, motion)
            p_, v_, w_ = g.state()
            ho[0][:], ho[1][:], ho[2][:] = p_, v_, w_
        hp, ho = ho, hp  # the next call starts from this call's result: the caller hands the output buffers back in

    e2e_step()  # warm
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    bytes_io = int(3 * pos.nbytes)
    rows = g.reduce(dem.RED_NUM_CONTACTS)
    cbar = rows / n
    rec = {"value": value, "ms_total": ms_total, "spheres": n, "contacts_per_sphere": cbar, "history_records": rows,
           "neighbor_list_rebuilds_in_timed_region": st1["rebuilds"] - st0["rebuilds"],
           "settle": {"timesteps": args.settle, "seconds": settle_s, "ke_per_sphere_before_J": ke0, "ke_per_sphere_after_J": ke1,
                      "max_speed_after_m_s": vmax},
           "e2e": {"value": n * S * e2e_steps / t_e2e, "unit": "sphere-steps/s", "h2d_bytes_per_step": bytes_io,
                   "d2h_bytes_per_step": bytes_io, "steps": e2e_steps},
           "clocks": sampler.result(), "wall_s_timed_region": t_wall}
    if cfg_id == 3:
        f, tq = g.mesh_wrench(0)
        rec["mesh_force"] = [float(x) for x in f]
        rec["mesh_facets"] = int(g.num_triangles)
        assert np.linalg.norm(f) > 0, "configs[3]: no sphere-facet contact in the timed region"
    if with_profile and motion is None:
        prof = g.step_profile(20)
        rec["kernel_ms_per_timestep"] = {k: v / 20 for k, v in prof.items()}
    return rec, g, scene, motion


def roofline_record(rec, n, world=1):
    hbm, hbm_src = peaks()
    cbar = rec["contacts_per_sphere"] or 6.0
    B_kernel = 160.0 + 32.0 * cbar   # narrowphase + force + integrate: state R+W 144, radius/id 16, history 32 per contact side
    B_step = 176.0 + 32.0 * cbar     # + 16 B cell key / permutation of the rebuild kernels
    out = {"bound": "hbm", "peak": hbm, "unit": "GB/s", "peak_source": hbm_src, "algorithmic_bytes_per_sphere_step": B_step,
           "whole_step_frac": B_step * (rec["value"] / world) / 1e9 / hbm}
    prof = rec.get("kernel_ms_per_timestep")
    if prof:
        dom = max(prof, key=prof.get)
        achieved = B_kernel * n / (prof[dom] * 1e-3) / 1e9
        traffic, src = measured_traffic(n, dom)
        out.update({"achieved": achieved, "frac": achieved / hbm, "traffic": traffic, "traffic_source": src, "kernel": dom,
                    "kernel_ms": prof[dom], "algorithmic_bytes_per_sphere": B_kernel})
    else:
        out.update({"achieved": B_step * (rec["value"] / world) / 1e9, "frac": out["whole_step_frac"], "traffic": None,
                    "kernel": "whole step, per GPU"})
    return out


def run_single(args):
    import torch
    from chrono_b200 import dem
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cfg_id = args.config
    S = args.substeps
    n = args.spheres or (WEAK_PER_GPU if cfg_id == 4 else DEFAULT_SPHERES_TOTAL[cfg_id])
    if cfg_id == 0:
        return run_config0(args, n)
    rec, g, scene, motion = measure_single(cfg_id, n, args, local)
    n = rec["spheres"]
    line = {
        "metric": "sphere-steps/sec", "value": rec["value"], "unit": "sphere-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_total"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg_id, n, n, S, 1, args.settle),
        "contacts_per_sphere": rec["contacts_per_sphere"], "history_records": rec["history_records"],
        "neighbor_list_rebuilds_in_timed_region": rec["neighbor_list_rebuilds_in_timed_region"], "settle": rec["settle"],
        "roofline": roofline_record(rec, n),
        "e2e": rec["e2e"],
        # conditional step graph: only k_step_begin + k_force_integrate launch in a step that does not rebuild
        "gpu_launches": int(args.steps * S * (2 if COND_GRAPH else KERNELS_PER_TIMESTEP + (5 if cfg_id == 3 else 0))),
        "clocks": rec["clocks"], "wall_s_timed_region": rec["wall_s_timed_region"],
    }
    for k in ("mesh_force", "mesh_facets"):
        if k in rec:
            line[k] = rec[k]
    prof = rec.get("kernel_ms_per_timestep")
    if prof:
        step_ms = sum(prof.values())
        line["kernel_ms_per_timestep"] = prof
        line["kernel_share"] = {k: v / step_ms for k, v in prof.items()}
    settled = g.state() if (args.cpu_steps > 0 and cfg_id in (1, 2)) else None
    # ---- the same bed, flowing: sheared at 1 m/s over its depth (rebuild cadence of a moving bed)
    if cfg_id == 1 and not args.no_flowing:
        pos, vel, om = g.state()
        z = pos[:, 2]
        vel = vel.copy()
        vel[:, 0] += 1.0 * (z - z.min()) / max(z.max() - z.min(), 1e-9)
        g.set_state(vel=vel)
        g.step(S, sync=False)
        g.sync()
        st0 = g.stats()
        ms = timed_single(g, S, max(3, args.steps // 2))
        st1 = g.stats()
        line["flowing"] = {"value": n * S * max(3, args.steps // 2) / (ms * 1e-3), "unit": "sphere-steps/s",
                           "what": "the settled bed given a linear shear profile, 1 m/s at the surface; timed right after",
                           "neighbor_list_rebuilds": st1["rebuilds"] - st0["rebuilds"], "timesteps": S * max(3, args.steps // 2),
                           "contacts_per_sphere": g.reduce(dem.RED_NUM_CONTACTS) / n}
    # ---- CPU baseline: the oracle port on the SETTLED bed (the GPU's positions / velocities; contact history starts empty)
    if args.cpu_steps > 0 and cfg_id in (1, 2):
        kw = physics(cfg_id)
        cpu_val, cpu_cores, cpu_t = cpu_oracle_run(scene, args.cpu_steps, state=settled, **kw)
        cpu_1t = cpu_oracle_run(scene, 1, threads=1, state=settled, **kw)[0]
        line["cpu_baseline"] = {"value": cpu_val, "unit": "sphere-steps/s", "cores": cpu_cores, "kind": "port",
                                "sample": "%d spheres x %d time steps from the settled state of the GPU run (same contacts per sphere)" % (n, args.cpu_steps),
                                "phase_seconds": cpu_t, "value_one_thread": cpu_1t}
    else:
        line["cpu_baseline"] = None
    g.close()
    del g
    # ---- the incumbent GPU implementation (unmodified Chrono::Dem CUDA) on the same box, same workload shape
    if cfg_id == 1 and not args.no_incumbent:
        line["incumbent_gpu"] = run_incumbent(n)
    # ---- like-for-like base of the weak-scaling lines: configs[4] (32 M spheres) on this one GPU
    if cfg_id == 1 and args.weak_base:
        a2 = argparse.Namespace(**vars(args))
        a2.steps, a2.warmup = max(3, args.steps // 4), 1
        r4 = measure_single(4, WEAK_PER_GPU, a2, local, with_profile=False)[0]
        line["weak_base"] = {"what": "configs[4] on ONE GPU: the per-GPU workload of the N > 1 lines", "value": r4["value"],
                             "unit": "sphere-steps/s", "spheres": r4["spheres"], "contacts_per_sphere": r4["contacts_per_sphere"],
                             "e2e": r4["e2e"]["value"], "settle": r4["settle"]}
    print(json.dumps(line))


def run_config0(args, n):
    """configs[0]: 10 k spheres, 1000 time steps, GPU and CPU both full length from the same initial packing."""
    import torch
    from chrono_b200 import dem, scenes
    scene = build_scene(0, n)
    kw = physics(0)
    g = scenes.make_gpu(scene, **kw)
    g.step(10)
    T = 1000
    ms = g.step_timed(T)
    cbar = g.reduce(dem.RED_NUM_CONTACTS) / n
    pos, vel, om = g.state()
    hp = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (pos, vel, om)]
    ho = [torch.empty(a.shape, dtype=torch.float64).pin_memory().numpy() for a in (pos, vel, om)]
    t0 = time.perf_counter()
    g.advance_host(hp[0], hp[1], hp[2], T, ho[0], ho[1], ho[2])
    t_e2e = time.perf_counter() - t0
    cpu_val, cores, cpu_t = cpu_oracle_run(scene, T, **kw)
    hbm, src = peaks()
    val = n * T / (ms * 1e-3)
    line = {"metric": "sphere-steps/sec", "value": val, "unit": "sphere-steps/s", "n_gpus": 1, "steps": 1, "warmup": 0,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(0, n, n, T, 1, 0), "contacts_per_sphere": cbar,
            "roofline": {"bound": "hbm", "achieved": (176 + 32 * cbar) * val / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": (176 + 32 * cbar) * val / 1e9 / hbm, "traffic": None, "kernel": "whole step",
                         "note": "10 k spheres = 313 warps on 148 SMs: launch-latency bound (9 launches per time step), not a roofline case"},
            "cpu_baseline": {"value": cpu_val, "unit": "sphere-steps/s", "cores": cores, "kind": "port",
                             "sample": "the full run: %d spheres x %d time steps" % (n, T), "phase_seconds": cpu_t},
            "e2e": {"value": n * T / t_e2e, "unit": "sphere-steps/s", "h2d_bytes_per_step": int(3 * pos.nbytes),
                    "d2h_bytes_per_step": int(3 * pos.nbytes), "steps": 1},
            "gpu_launches": T * KERNELS_PER_TIMESTEP}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# N > 1: slabs
# ---------------------------------------------------------------------------------------------------------------------
def slab_scene(cfg_id, n_total, world, rank):
    """(local pos, radius, ids, lo, hi, global scene dict) of rank's slab."""
    from chrono_b200 import scenes, slab
    if cfg_id in (1, 4):
        sc = scenes.slab_lattice_scene(n_total, world=world, rank=rank, radius=RADIUS, layers=BED_LAYERS, precompress=True)
        return sc["pos"], sc["radius"], sc["ids"], sc["lo"], sc["hi"], sc
    if cfg_id == 2:
        sc = scenes.slab_lattice_scene(n_total, world=world, rank=rank, radius=RADIUS, sep_factor=1.75, polydisperse=(0.8, 1.2))
        return sc["pos"], sc["radius"], sc["ids"], sc["lo"], sc["hi"], sc
    sc = build_scene(cfg_id, n_total)  # drum: every rank generates the packing and keeps its stretch of the axis
    x = sc["pos"][:, 0]
    bounds = slab.slab_bounds(x, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = np.nonzero((x >= lo) & (x < hi))[0]
    sc["n_total"] = sc["n"]
    sc["rmax"] = float(sc["radius"].max())
    return sc["pos"][mine], sc["radius"][mine], mine.astype(np.uint32), lo, hi, sc


def measure_slabs(cfg_id, n_total, args, rank, world, local, dist, torch):
    # the engine runs on the stream that is current when it is created: give it a stream of its own instead of the legacy
    # default stream (graph launches into stream 0 serialise with every other blocking stream of the process)
    with torch.cuda.stream(torch.cuda.Stream()):
        return _measure_slabs(cfg_id, n_total, args, rank, world, local, dist, torch)


def _measure_slabs(cfg_id, n_total, args, rank, world, local, dist, torch):
    from chrono_b200 import dem, scenes, slab
    S = args.substeps
    pos, rad, ids, lo, hi, sc = slab_scene(cfg_id, n_total, world, rank)
    n_total = sc["n_total"]
    kw = physics(cfg_id)
    mat = kw.pop("mat", None) or scenes.settling_material()
    kw.pop("dt", None)
    cfg = dem.config(device=local, dt=DT, bins=sc["bins"], mat_sphere=dem.material(**mat), mat_wall=dem.material(**mat),
                     mat_mesh=dem.material(**mat), mass_coef=scenes.MASS_COEF, wall_mass=1.0, integrator=dem.CENTERED_DIFFERENCE,
                     history_slots=kw.pop("history_slots", 16), neighbor_slots=kw.pop("neighbor_slots", 0),
                     verlet_skin=args.skin * RADIUS if args.skin > 0 else -1.0, **kw)
    g, backend = slab.make_engine_slab(cfg, sc["walls"], pos, rad, ids, capacity=int(1.12 * len(ids)) + 65536,
                                       rmax_global=sc["rmax"], meshes=sc.get("meshes"))
    del pos, rad
    drv = slab.SlabDriver(backend, rank, world, lo, hi, lag=1)
    drv.rebuild()
    if not args.nccl_halo:
        drv.enable_p2p(lag=args.slab_lag)
    motion = DrumMotion(sc) if cfg_id == 3 else None

    def advance_slab(k):
        if motion is None:
            drv.step(k)
        else:
            for _ in range(k):
                motion.apply(g)
                drv.step(1)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    t0 = time.perf_counter()
    advance_slab(args.settle)
    drv.drain()
    g.sync()
    barrier()
    settle_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        advance_slab(S)
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
        advance_slab(S)
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
    value = n_total * nsteps / (ms_total * 1e-3)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)
    st1 = dict(drv.stats)

    # ---- e2e: every bench step the owned state comes from pinned host memory and goes back to it
    sid, p0, v0, w0 = backend.export_owned()
    n_own = len(sid)
    bytes_io = int(3 * p0.nbytes)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        drv.import_owned(p0, v0, w0)  # H2D of pos + vel + omega of the owned spheres, then the slab rebuild it implies
        advance_slab(S)
        drv.drain()
        sid, p0, v0, w0 = backend.export_owned()  # D2H of the owned spheres (ids + pos + vel + omega)
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_total * S * e2e_steps / float(te.item())

    halo = torch.tensor([st1["halo_bytes"] - st0["halo_bytes"], st1["migrated"] - st0["migrated"],
                         st1["rebuilds"] - st0["rebuilds"], n_own, rebuild_s_timed, bytes_io], device="cuda", dtype=torch.float64)
    allh = [torch.zeros_like(halo) for _ in range(world)]
    dist.all_gather(allh, halo)
    red = torch.tensor([g.reduce(dem.RED_NUM_CONTACTS), g.reduce(dem.RED_KE)], device="cuda", dtype=torch.float64)
    dist.all_reduce(red)
    vmax = torch.tensor([g.reduce(dem.RED_MAX_SPEED)], device="cuda", dtype=torch.float64)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    rec = {"value": value, "ms_total": ms_total, "spheres": n_total, "contacts_per_sphere": float(red[0]) / n_total,
           "settle": {"timesteps": args.settle, "seconds": settle_s, "ke_per_sphere_after_J": float(red[1]) / n_total,
                      "max_speed_after_m_s": float(vmax.item())},
           "e2e": {"value": e2e_value, "unit": "sphere-steps/s", "h2d_bytes_per_step": int(sum(float(h[5]) for h in allh)),
                   "d2h_bytes_per_step": int(sum(float(h[5]) for h in allh)), "steps": e2e_steps,
                   "note": "per bench step: owned pos / vel / omega H2D (dem_b200_import_owned) + slab rebuild + S time steps + D2H "
                           "(dem_b200_export_owned); bytes summed over the ranks"},
           "halo": {"bytes_received_per_rank_per_timestep": [float(h[0]) / nsteps for h in allh],
                    "nvlink_GBps_busiest_rank": max(float(h[0]) for h in allh) / (ms_total * 1e-3) / 1e9,
                    "nvlink_peak_GBps_per_direction": 900.0,
                    "migrated_spheres": [int(h[1]) for h in allh], "rebuilds": int(allh[0][2]),
                    "owned_spheres": [int(h[3]) for h in allh],
                    "rebuild_ms_total_in_timed_region": [1e3 * float(h[4]) for h in allh],
                    "ms_per_timestep_outside_rebuilds": (ms_total - 1e3 * max(float(h[4]) for h in allh)) / nsteps,
                    "transport": "NCCL send/recv + 4-byte all-reduce per time step" if args.nccl_halo else
                    "NVLink peer stores from the pack kernel (CUDA IPC), device-side flags and vote; no collective per time step"},
           "clocks": sampler.result() if rank == 0 else None, "wall_s_timed_region": t_wall, "host_enqueue_s": t_enqueue}
    if cfg_id == 3:
        w = torch.tensor(g.mesh_wrench(0)[0], device="cuda", dtype=torch.float64)
        dist.all_reduce(w)
        rec["mesh_force"] = [float(x) for x in w.cpu().numpy()]
        rec["mesh_facets"] = int(g.num_triangles)
        assert float(torch.linalg.norm(w)) > 0, "configs[3]: no sphere-facet contact in the timed region"
    g.close()
    del g, backend, drv
    torch.cuda.empty_cache()
    return rec


def run_slabs(args):
    """N > 1: slab domain decomposition (chrono_b200/slab.py).  One process per GPU; per time step: ghost halo (pos, v, omega
    of the spheres within 2 r_max + skin of a slab face) by NVLink peer stores, the same step graph as on one GPU, a
    device-side vote on "rebuild now?"; spheres migrate at rebuilds."""
    import torch
    import torch.distributed as dist
    from chrono_b200 import slab

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_to_gpu_numa_node(local)
    # NCCL writes its version banner (NCCL_DEBUG >= VERSION) to stdout: send its log to stderr, rank 0 prints ONE JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg_id = args.config
    S = args.substeps
    if args.spheres:
        n_total = args.spheres * world
    elif cfg_id == 4:
        n_total = WEAK_PER_GPU * world
    elif cfg_id == 1:
        n_total = 1000000 * world
    else:
        n_total = DEFAULT_SPHERES_TOTAL[cfg_id]
    # ---- driver-side proof of the N-GPU path: a slab run against the same job on one GPU, bit for bit, before anything is timed
    parity = None
    if not args.no_parity:
        parity = slab.slab_parity_check(rank, world, local, n=args.parity_spheres, steps=args.parity_steps, p2p=not args.nccl_halo)
        if not (parity["bit_identical"] and parity["contacts_equal"]):
            if rank == 0:
                print(json.dumps({"error": "slab run differs from the single-GPU run", "parity": parity}))
            dist.destroy_process_group()
            sys.exit(3)
    rec = measure_slabs(cfg_id, n_total, args, rank, world, local, dist, torch)
    n_total = rec["spheres"]
    strong = None
    if cfg_id == 4 and not args.no_strong and not args.spheres and STRONG_TOTAL != WEAK_PER_GPU * world:
        a2 = argparse.Namespace(**vars(args))
        a2.steps, a2.warmup = max(3, args.steps // 2), 2
        r2 = measure_slabs(4, STRONG_TOTAL, a2, rank, world, local, dist, torch)
        strong = {"what": "strong scaling: %d spheres in total over %d GPUs" % (r2["spheres"], world),
                  "value": r2["value"], "unit": "sphere-steps/s", "spheres_total": r2["spheres"], "spheres_per_gpu": r2["spheres"] // world,
                  "contacts_per_sphere": r2["contacts_per_sphere"], "ms_per_timestep": r2["ms_total"] / (a2.steps * S),
                  "e2e": r2["e2e"]["value"], "halo_GBps_busiest_rank": r2["halo"]["nvlink_GBps_busiest_rank"],
                  "rebuilds": r2["halo"]["rebuilds"]}
    hbm, hbm_src = peaks()
    if rank == 0:
        line = {
            "metric": "sphere-steps/sec", "value": rec["value"], "unit": "sphere-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_total"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg_id, n_total // world, n_total, S, world, args.settle),
            "contacts_per_sphere": rec["contacts_per_sphere"], "settle": rec["settle"],
            "parity": parity,
            "scaling_note": "weak scaling: every GPU holds %d spheres; the like-for-like one-GPU figure is `weak_base` of the "
                            "`bench.py --gpus 1` line (same generator, same per-GPU size), not its configs[1] headline" % (n_total // world),
            "roofline": roofline_record(rec, n_total // world, world),
            "halo": rec["halo"], "cpu_baseline": None, "e2e": rec["e2e"],
            "gpu_launches": int(args.steps * S * (KERNELS_PER_TIMESTEP_SLAB + (5 if cfg_id == 3 else 0))),
            "clocks": rec["clocks"], "wall_s_timed_region": rec["wall_s_timed_region"], "host_enqueue_s": rec["host_enqueue_s"],
            "cpu_affinity_rank0": affinity,
        }
        if strong:
            line["strong"] = strong
        elif cfg_id == 4 and not args.spheres and STRONG_TOTAL == WEAK_PER_GPU * world:
            line["strong"] = {"what": "strong scaling: %d spheres in total over %d GPUs -- at this N the weak-scaling line above IS that run" % (n_total, world),
                              "value": rec["value"], "unit": "sphere-steps/s", "spheres_total": n_total}
        for k in ("mesh_force", "mesh_facets"):
            if k in rec:
                line[k] = rec[k]
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=-1, choices=[-1, 0, 1, 2, 3, 4], help="BASELINE.json configs[i]; default 1 on one GPU, 4 on several")
    ap.add_argument("--spheres", type=int, default=0, help="spheres per GPU (0 = the size the config names)")
    ap.add_argument("--substeps", type=int, default=100, help="DEM time steps per bench step")
    ap.add_argument("--settle", type=int, default=-1, help="untimed time steps that settle the bed before warm-up (default: 3000; configs[2] 10000)")
    ap.add_argument("--ref-substeps", type=int, default=2, help="time steps per bench step of the reference arm")
    ap.add_argument("--cpu-steps", type=int, default=4, help="time steps of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-flowing", action="store_true", help="N = 1: skip the flowing-bed record")
    ap.add_argument("--no-incumbent", action="store_true", help="N = 1: skip the incumbent Chrono::Dem CUDA run")
    ap.add_argument("--weak-base", type=int, default=1, help="N = 1, config 1: also time configs[4] (32 M spheres) on this GPU (0 = skip)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1, config 4: skip the strong-scaling record")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the slab-vs-single-GPU bit-identity run")
    ap.add_argument("--parity-spheres", type=int, default=200000)
    ap.add_argument("--parity-steps", type=int, default=300)
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: NCCL send/recv halo + all-reduce vote instead of P2P stores")
    ap.add_argument("--slab-lag", type=int, default=3, help="N > 1: steps between casting the rebuild vote and acting on it")
    ap.add_argument("--skin", type=float, default=0.0, help="fixed Verlet skin in sphere radii (0 = engine default: adaptive 0.25 - 0.5 on one GPU, 0.35 on slabs)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config < 0:
        args.config = 1 if world == 1 else 4
    if args.settle < 0:
        args.settle = 10000 if args.config == 2 else (0 if args.config == 0 else 3000)
    if args.impl == "reference":
        run_reference(args)
    elif world > 1:
        run_slabs(args)
    else:
        run_single(args)


if __name__ == "__main__":
    main()

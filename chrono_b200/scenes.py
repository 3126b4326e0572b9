"""Deterministic synthetic scenes: HCP lattice fill of a five-wall box container.

Geometry follows the reference utilities so that the CPU oracle and the GPU engine can be fed
bit-identical inputs:
  * lattice      ChHCPSampler::Sample        src/chrono/utils/ChUtilsSamplers.h:540-568
  * container    utils::AddBoxContainer      src/chrono/utils/ChUtilsCreators.cpp:559-615
  * material     btest_MCORE_settling.cpp:80-112 (Y=2e6, mu=0.4, cr=0.4, rho=2000, R=0.02)
"""
import math

import numpy as np


def hcp_points(lo, hi, sep):
    """HCP lattice points p with lo <= p <= hi (ChHCPSampler, box volume).  Returned in the reference's
    generation order: z layers, then y rows, then x."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    dx = sep
    dy = sep * (math.sqrt(3.0) / 2)
    dz = sep * math.sqrt(2.0 / 3.0)
    size = hi - lo
    nx = int(size[0] / dx) + 1
    ny = int(size[1] / dy) + 1
    nz = int(size[2] / dz) + 1
    k = np.arange(nz)[:, None, None]
    j = np.arange(ny)[None, :, None]
    i = np.arange(nx)[None, None, :]
    offy = np.where(k % 2 == 0, 0.0, dy / 3)
    offx = np.where((j + k) % 2 == 0, 0.0, dx / 2)
    x = lo[0] + offx + i * dx + 0 * k
    y = lo[1] + offy + j * dy + 0 * i
    z = lo[2] + k * dz + 0 * j + 0 * i
    p = np.stack([np.broadcast_to(x, (nz, ny, nx)), np.broadcast_to(y, (nz, ny, nx)),
                  np.broadcast_to(z, (nz, ny, nx))], axis=-1).reshape(-1, 3)
    ok = np.all((p >= lo - 1e-12) & (p <= hi + 1e-12), axis=1)
    return np.ascontiguousarray(p[ok])


def box_container(size, thickness, center=(0.0, 0.0, 0.0), faces=(2, 2, -1)):
    """(pos, hdims) of the wall boxes, in AddBoxContainer order: Z-, Z+, X-, X+, Y-, Y+ (only the requested
    faces; -1: negative side, +1: positive side, 2: both)."""
    size = np.asarray(size, dtype=np.float64)
    c = np.asarray(center, dtype=np.float64)
    hs = size / 2
    ht = thickness / 2
    walls = []

    def add(axis, sign):
        p = np.zeros(3)
        p[axis] = sign * (hs[axis] + ht)
        dims = size.copy()
        dims[axis] = thickness
        walls.append((c + p, dims / 2))

    order = [(2, faces[2]), (0, faces[0]), (1, faces[1])]
    for axis, f in order:
        if f in (-1, 2):
            add(axis, -1.0)
        if f in (1, 2):
            add(axis, +1.0)
    return walls


def settling_scene(n_target, radius=0.02, box_xy=None, seed=12345, jitter=0.005, sep_factor=2.02,
                   polydisperse=None, wall_thickness=0.2, bin_factor=2.0, headroom=1.25):
    """N spheres on a jittered HCP lattice inside an open-top box.

    Returns dict(pos, radius, box_size, walls, bins, n).  box_xy: (Lx, Ly) in metres, chosen from n_target for
    a bed roughly 0.26 x as deep as wide when None.  bins: broadphase resolution with bin edge >= bin_factor * Rmax
    in every direction over the (inflated) wall bounding box."""
    rng = np.random.default_rng(seed)
    sep = sep_factor * radius * (1.2 if polydisperse else 1.0)
    dxv = sep ** 3 / math.sqrt(2.0)  # volume per HCP site
    if box_xy is None:
        L = (n_target * dxv / 0.26) ** (1.0 / 3.0)
        box_xy = (L, L)
    Lx, Ly = box_xy
    rmax = radius * (polydisperse[1] if polydisperse else 1.0)
    lo = np.array([-Lx / 2 + rmax * 1.01, -Ly / 2 + rmax * 1.01, rmax * 1.01])
    # fill height needed for n_target sites
    per_layer = max(1, int((Lx - 2 * rmax) / sep) + 1) * max(1, int((Ly - 2 * rmax) / (sep * math.sqrt(3) / 2)) + 1)
    layers = int(math.ceil(n_target / per_layer)) + 2
    hi = np.array([Lx / 2 - rmax * 1.01, Ly / 2 - rmax * 1.01, lo[2] + layers * sep * math.sqrt(2.0 / 3.0)])
    pts = hcp_points(lo, hi, sep)
    if len(pts) < n_target:
        raise ValueError("box too small for n_target")
    pts = pts[:n_target]
    pts = pts + rng.uniform(-jitter * radius, jitter * radius, size=pts.shape)
    if polydisperse:
        rad = radius * rng.uniform(polydisperse[0], polydisperse[1], size=n_target)
    else:
        rad = np.full(n_target, radius)
    Lz = max(pts[:, 2].max() + 2 * rmax, 0.25 * Lx) * headroom
    size = np.array([Lx, Ly, Lz])
    walls = box_container(size, wall_thickness, center=(0.0, 0.0, Lz / 2), faces=(2, 2, -1))
    # bounding box of the walls, inflated like ChBroadphase::DetermineBoundingBox, sets the grid extent
    wmin = np.min([p - h for p, h in walls], axis=0)
    wmax = np.max([p + h for p, h in walls], axis=0)
    ext = (wmax - wmin) * 1.002
    bins = np.maximum(1, np.floor(ext / (bin_factor * rmax * 1.0001)).astype(np.int64))
    return dict(pos=np.ascontiguousarray(pts), radius=rad, box_size=size, walls=walls,
                bins=tuple(int(b) for b in bins), n=n_target)

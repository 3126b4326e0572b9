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


def quat_from_axis_angle(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    return np.concatenate([[math.cos(angle / 2)], math.sin(angle / 2) * a])


def quat_rotate(v, q):
    """Rotate (src/chrono/multicore_math/real4.cpp:158-161) for an array of vectors and one quaternion (w,x,y,z)."""
    v = np.asarray(v, dtype=np.float64)
    u = np.asarray(q[1:], dtype=np.float64)
    t = 2 * np.cross(u, v)
    return v + q[0] * t + np.cross(u, t)


def heightfield_mesh(x0, x1, y0, y1, nx, ny, z_fn, flip=False):
    """Triangulated height field z = z_fn(x, y) over [x0,x1] x [y0,y1] with nx x ny cells: (2 nx ny, 9) vertices.
    Winding: (B-A) x (C-A) points to +z (the side Multicore's one-sided triangle_sphere collides on) unless flip."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    Z = z_fn(X, Y)
    P = np.stack([X, Y, Z], axis=-1)
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t1 = np.concatenate([a, b, c], axis=-1).reshape(-1, 9)
    t2 = np.concatenate([a, c, d], axis=-1).reshape(-1, 9)
    tri = np.concatenate([t1, t2], axis=0)
    if flip:
        tri = tri.reshape(-1, 3, 3)[:, [0, 2, 1], :].reshape(-1, 9)
    return np.ascontiguousarray(tri)


def mesh_to_body_frame(tri_world, pos, rot):
    """World-frame triangles -> frame of a body at (pos, rot): v_loc = RotateT(v - pos, rot)."""
    q = np.asarray(rot, dtype=np.float64)
    qc = np.concatenate([q[:1], -q[1:]])
    v = np.asarray(tri_world, dtype=np.float64).reshape(-1, 3) - np.asarray(pos, dtype=np.float64)
    return np.ascontiguousarray(quat_rotate(v, qc).reshape(-1, 9))


def cylinder_drum_mesh(radius, length, n_seg, axis=1, caps=True, n_ax=1):
    """Closed cylinder (the drum of BASELINE configs[3]) seen from inside: normals point to the axis.  Axis y (axis=1)
    by default, centred at the origin.  n_seg segments around x n_ax along the axis, 2 triangles each, + 2 n_seg cap
    triangles (fans)."""
    th = np.linspace(0.0, 2 * math.pi, n_seg + 1)
    c, s_ = np.cos(th), np.sin(th)
    h = length / 2
    ys = np.linspace(-h, h, n_ax + 1)
    tris = []

    def P(i, yy):
        return [radius * c[i], yy, radius * s_[i]]

    for i in range(n_seg):
        for k in range(n_ax):
            a, b, cc, d = P(i, ys[k]), P(i + 1, ys[k]), P(i + 1, ys[k + 1]), P(i, ys[k + 1])
            # inward normal: for axis y the outward radial is (cos, 0, sin); the winding is fixed by the test below
            tris.append(a + cc + b)
            tris.append(a + d + cc)
        if caps:
            tris.append([0, -h, 0] + P(i, -h) + P(i + 1, -h))   # normal +y (into the drum) checked below
            tris.append([0, h, 0] + P(i + 1, h) + P(i, h))      # normal -y
    tri = np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)
    # make every normal point towards the drum centre
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    cen = tri.mean(axis=1)
    wrong = np.einsum("ij,ij->i", n, -cen) < 0
    tri[wrong] = tri[wrong][:, [0, 2, 1], :]
    if axis != 1:
        perm = {0: [1, 0, 2], 2: [0, 2, 1]}[axis]
        tri = tri[:, :, perm]
        tri = tri[:, [0, 2, 1], :]  # a coordinate swap mirrors the winding
    return np.ascontiguousarray(tri.reshape(-1, 9))

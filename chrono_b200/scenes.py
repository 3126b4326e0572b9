"""Deterministic synthetic scenes: HCP lattice fill of a five-wall box container.

Geometry follows the reference utilities so that the CPU oracle and the GPU engine can be fed
bit-identical inputs:
  * lattice      ChHCPSampler::Sample        src/chrono/utils/ChUtilsSamplers.h:540-568
  * container    utils::AddBoxContainer      src/chrono/utils/ChUtilsCreators.cpp:559-615
  * material     btest_MCORE_settling.cpp:80-112 (Y=2e6, mu=0.4, cr=0.4, rho=2000, R=0.02)
"""
import math

import numpy as np

RHO = 2000.0
MASS_COEF = 4.0 / 3.0 * np.pi * RHO


def settling_material(mu=0.4, cr=0.4, young=2e6, mu_roll=0.0, mu_spin=0.0, adhesion=0.0):
    """btest_MCORE_settling.cpp:80-92"""
    return dict(young=young, poisson=0.3, mu_s=mu, mu_roll=mu_roll, mu_spin=mu_spin, cr=cr, adhesion=adhesion)


def make_gpu(scene, dt=1e-3, mat=None, wall_mat=None, mesh_mat=None, wall_mass=1.0, vel=None, omega=None, gravity=(0, 0, -9.81),
             integrator=None, history_slots=16, device=0, stream=None, **model):
    """The scene in the CUDA engine, through its C ABI (chrono_b200/dem.py).  No CPU path: raises when the library is missing.
    stream: raw cudaStream_t handle the engine runs on instead of its own stream (so that the caller's events time it)."""
    from . import dem
    mat = mat or settling_material()
    cfg = dem.config(device=device, dt=dt, bins=scene["bins"], gravity=gravity, mat_sphere=dem.material(**mat),
                     mat_wall=dem.material(**(wall_mat or mat)), mat_mesh=dem.material(**(mesh_mat or mat)),
                     mass_coef=MASS_COEF, wall_mass=wall_mass,
                     integrator=dem.CENTERED_DIFFERENCE if integrator is None else integrator,
                     history_slots=history_slots, **model)
    g = dem.DemSystem(cfg)
    if stream is not None:
        import ctypes
        g._ck(g.L.dem_b200_set_stream(g.h, ctypes.c_void_p(stream)))
    for p, h in scene["walls"]:
        g.add_box_wall(p, h)
    for c, rb in scene.get("balls", []):
        g.add_sphere_wall(c, rb, spheres_outside=True)
    for M in scene.get("meshes", []):
        m = g.add_mesh(M["tri"], M.get("mass", 1.0))
        g.set_mesh_motion(m, M.get("pos"), M.get("rot"), M.get("vel"), M.get("omega"))
    g.set_spheres(scene["pos"], scene["radius"], vel=vel, omega=omega, fixed=scene.get("fixed"))
    g.initialize()
    return g


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


# ---------------------------------------------------------------------------------------------------------------------
# large scenes generated per slab (BASELINE configs[4]): no rank ever holds the whole packing
# ---------------------------------------------------------------------------------------------------------------------
def _hash_uniform(gid, stream):
    """Counter-based uniform numbers in [-1, 1): splitmix64 of (sphere id, stream).  The jitter of a sphere depends on its
    global id only, so every partition of the lattice generates bit-identical spheres."""
    z = gid.astype(np.uint64) * np.uint64(4) + np.uint64(stream) + np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def hcp_layer_heights(nz, radius, sep, young=2e6, poisson=0.3, g=9.81):
    """Centre heights of nz HCP layers (in-plane spacing sep, laterally confined) in STATIC EQUILIBRIUM under gravity with the
    Hertz law of the engine (ChIterativeSolverMulticoreSMC.cpp:277-290: F = kn delta, kn = 4/3 E* sqrt(R* delta)): every sphere
    rests on three spheres of the layer below (horizontal offset sep / sqrt(3)), the bottom layer on the floor at z = 0.
    A bed generated at these heights starts at rest instead of collapsing by several per cent of its depth."""
    E = 1.0 / (2.0 * (1.0 - poisson * poisson) / young)        # composite of two equal materials (make_comp, float there)
    m = MASS_COEF * radius ** 3
    h = sep / math.sqrt(3.0)
    dz0 = sep * math.sqrt(2.0 / 3.0)
    n_above = np.arange(nz - 1, 0, -1, dtype=np.float64)       # layers resting on interface q = 0 .. nz-2
    kH = 4.0 / 3.0 * E * math.sqrt(0.5 * radius)
    lo, hi = np.zeros(nz - 1), np.full(nz - 1, dz0)
    for _ in range(80):                                          # bisection on the layer spacing: the support grows as dz shrinks
        dz = 0.5 * (lo + hi)
        dist = np.sqrt(h * h + dz * dz)
        delta = np.maximum(2.0 * radius - dist, 0.0)
        support = 3.0 * kH * delta ** 1.5 * dz / dist
        too_low = support > n_above * m * g
        lo = np.where(too_low, dz, lo)
        hi = np.where(too_low, hi, dz)
    dz = 0.5 * (lo + hi)
    dw = (nz * m * g / (4.0 / 3.0 * E * math.sqrt(radius))) ** (2.0 / 3.0)  # floor contact: face of a box, R* = R
    return np.concatenate([[radius - dw], radius - dw + np.cumsum(dz)])


def slab_lattice_scene(n_total, world=1, rank=0, radius=0.02, jitter=0.005, sep_factor=2.0, polydisperse=None,
                       wall_thickness=0.2, bin_factor=2.0, headroom=1.25, cross_section_of=None, layers=None, precompress=False):
    """The settling_scene packing for `world` slabs along x, of which only slab `rank` is generated.

    The box is `world` times as long (x) as the box of an n_total / world packing is wide, same width (y) and bed depth,
    made of complete HCP layers (so the sphere count is n_actual ~ n_total, returned).  Slab faces lie between lattice
    columns; global id = lattice index in generation order (z layers, y rows, x).  layers: fixed bed depth in HCP layers (the
    box widens instead of deepening as n_total grows: equal work per sphere at every size); precompress: layer heights in
    static equilibrium (hcp_layer_heights; monodisperse only).  Returns dict(pos, radius, ids, lo, hi, n_total, box_size,
    walls, bins) with pos / radius / ids of slab `rank` only."""
    sep = sep_factor * radius * (1.2 if polydisperse else 1.0)
    rmax = radius * (polydisperse[1] if polydisperse else 1.0)
    site = sep ** 3 / math.sqrt(2.0)
    n_one = cross_section_of or max(1, n_total // world)
    dx, dy, dz = sep, sep * (math.sqrt(3.0) / 2), sep * math.sqrt(2.0 / 3.0)
    if layers:
        L = math.sqrt(n_one * dx * dy / layers) + 2.02 * rmax   # square footprint holding n_one / layers sites
    else:
        L = (n_one * site / 0.26) ** (1.0 / 3.0)
    nx1 = max(1, int(round((L - 2.02 * rmax) / dx)))   # lattice columns per slab
    nx = nx1 * world
    ny = max(1, int(round((L - 2.02 * rmax) / dy)))
    nz = int(layers) if layers else max(1, int(round(n_total / float(nx * ny))))
    # the box holds the lattice exactly: odd rows are shifted by dx / 2, odd layers by dy / 3
    Lx, Ly = (nx - 1) * dx + dx / 2 + 2.02 * rmax, (ny - 1) * dy + dy / 3 + 2.02 * rmax
    x0, y0, z0 = -Lx / 2 + 1.01 * rmax, -Ly / 2 + 1.01 * rmax, 1.01 * rmax
    i0, i1 = rank * nx1, (rank + 1) * nx1
    k = np.arange(nz, dtype=np.int64)[:, None, None]
    j = np.arange(ny, dtype=np.int64)[None, :, None]
    i = np.arange(i0, i1, dtype=np.int64)[None, None, :]
    gid = ((k * ny + j) * nx + i).reshape(-1)
    offy = np.where(k % 2 == 0, 0.0, dy / 3)
    offx = np.where((j + k) % 2 == 0, 0.0, dx / 2)
    shape = (nz, ny, i1 - i0)
    pos = np.empty((gid.size, 3))
    pos[:, 0] = np.broadcast_to(x0 + offx + i * dx, shape).reshape(-1)
    pos[:, 1] = np.broadcast_to(y0 + offy + j * dy + 0.0 * i, shape).reshape(-1)
    if precompress and not polydisperse:
        zk = hcp_layer_heights(nz, radius, sep)
    else:
        zk = z0 + np.arange(nz) * dz
    pos[:, 2] = np.broadcast_to(zk[:, None, None] + 0.0 * j + 0.0 * i, shape).reshape(-1)
    for c in range(3):
        pos[:, c] += (jitter * radius) * _hash_uniform(gid, c)
    if polydisperse:
        u = 0.5 * (_hash_uniform(gid, 3) + 1.0)
        rad = radius * (polydisperse[0] + (polydisperse[1] - polydisperse[0]) * u)
    else:
        rad = np.full(gid.size, radius)
    ztop = zk[-1] + jitter * radius
    Lz = max(ztop + 2 * rmax, 0.25 * Ly) * headroom
    size = np.array([Lx, Ly, Lz])
    walls = box_container(size, wall_thickness, center=(0.0, 0.0, Lz / 2), faces=(2, 2, -1))
    wmin = np.min([p - h for p, h in walls], axis=0)
    wmax = np.max([p + h for p, h in walls], axis=0)
    bins = np.maximum(1, np.floor((wmax - wmin) * 1.002 / (bin_factor * rmax * 1.0001)).astype(np.int64))
    # slab faces half-way between the last column of one slab and the first of the next (both row offsets considered)
    face = lambda c: x0 + c * dx - dx / 4
    lo = -np.inf if rank == 0 else face(i0)
    hi = np.inf if rank == world - 1 else face(i1)
    return dict(pos=pos, radius=rad, ids=gid.astype(np.uint32), lo=lo, hi=hi, n_total=int(nx) * int(ny) * int(nz), n=int(gid.size),
                box_size=size, walls=walls, bins=tuple(int(b) for b in bins), rmax=float(rmax))


def drum_scene(n_target, radius=0.02, seed=7, fill=0.35, facet_edge=4.0, jitter=0.005, sep_factor=2.0, aspect=4.0):
    """BASELINE configs[3]: spheres resting in the lower part of a closed drum (triangle mesh, axis x) that rotates about
    its axis.  The drum is long (length = aspect * diameter) so that slabs ALONG THE AXIS (x, the slab direction of the
    engine) each own a stretch of drum wall; facets are about facet_edge sphere radii wide.
    Returns a scene dict whose `meshes[0]` is the drum; the spheres start on an HCP lattice clipped to the drum interior
    below the fill level, a hair above the wall, so that they are on the wall within a few hundred steps."""
    rng = np.random.default_rng(seed)
    sep = sep_factor * radius
    site = sep ** 3 / math.sqrt(2.0)
    # filled segment area fraction of a circle at relative fill height f (of the diameter)
    th = 2 * math.acos(1 - 2 * fill)
    seg_frac = (th - math.sin(th)) / (2 * math.pi)
    # volume needed -> drum radius: n * site = seg_frac * pi Rd^2 * (aspect * 2 Rd)
    Rd = (n_target * site / (seg_frac * math.pi * 2 * aspect)) ** (1.0 / 3.0) * 1.02
    for _ in range(40):
        Ld = 2 * aspect * Rd
        lo = np.array([-Rd, -Ld / 2 + 1.2 * radius, -Rd])
        hi = np.array([Rd, Ld / 2 - 1.2 * radius, -Rd + 2 * Rd * fill])
        pts = hcp_points(lo, hi, sep)
        rr = np.sqrt(pts[:, 0] ** 2 + pts[:, 2] ** 2)
        pts = pts[rr < Rd - 1.3 * radius]
        if len(pts) >= n_target:
            break
        Rd *= 1.02
    else:
        raise ValueError("drum_scene: could not place %d spheres" % n_target)
    pts = pts[np.argsort(pts[:, 2], kind="stable")][:n_target]  # the lowest sites: a bed with a level surface
    pts = pts + rng.uniform(-jitter * radius, jitter * radius, size=pts.shape)
    n_seg = max(24, int(2 * math.pi * Rd / (facet_edge * radius)))
    n_ax = max(1, int(round(Ld / (facet_edge * radius))))
    tri = cylinder_drum_mesh(Rd, Ld, n_seg, axis=0, caps=False, n_ax=n_ax)  # the end caps are two box walls (fan facets would be metres long)
    pts = np.ascontiguousarray(pts[:, [1, 0, 2]])  # the lattice was laid out along y: turn the drum axis to x
    ext = np.array([Ld, 2 * Rd, 2 * Rd]) * 1.05
    bins = np.maximum(1, np.floor(ext / (2.0 * radius * 1.0001)).astype(np.int64))
    mesh = dict(tri=tri, pos=np.zeros(3), rot=np.array([1.0, 0, 0, 0]), vel=np.zeros(3), omega=np.zeros(3), mass=1000.0)
    t = 0.2
    walls = [(np.array([s_ * (Ld / 2 + t / 2), 0.0, 0.0]), np.array([t / 2, 1.05 * Rd, 1.05 * Rd])) for s_ in (-1.0, 1.0)]
    return dict(pos=np.ascontiguousarray(pts), radius=np.full(n_target, radius), walls=walls, bins=tuple(int(b) for b in bins),
                n=n_target, meshes=[mesh], drum_radius=Rd, drum_length=Ld, box_size=ext)

"""Shared builders: the same scene in the CPU oracle and in the CUDA engine (through its C ABI)."""
import numpy as np

from oracle import pyoracle as po

from chrono_b200.scenes import MASS_COEF, RHO, settling_material  # noqa: F401


def sphere_mass(radius, mass_coef=MASS_COEF):
    """Same expression (and rounding sequence) as sphere_mass() in chrono_b200/csrc/dem_kernels.cuh."""
    r = np.asarray(radius, dtype=np.float64)
    return mass_coef * (r * r * r)


def make_oracle(scene, dt=1e-3, mat=None, wall_mat=None, mesh_mat=None, wall_mass=1.0, vel=None, omega=None, gravity=(0, 0, -9.81),
                num_threads=0, history_slots=None, integrator=None, **model):
    mat = mat or settling_material()
    s = po.make_settings(dt=dt, bins=scene["bins"], gravity=gravity, num_threads=num_threads, **model)
    o = po.Oracle(s)
    m_s = o.add_material(po.make_material(**mat))
    m_w = o.add_material(po.make_material(**(wall_mat or mat)))
    # body 0: the container (fixed), boxes first -> shapes 0..nW-1
    if scene["walls"]:
        o.add_body(wall_mass, (1, 1, 1), (0, 0, 0), fixed=True)
        for p, h in scene["walls"]:
            o.add_box(0, m_w, p, h)
    # ball boundaries (Chrono::Dem CreateBCSphere): fixed sphere bodies right behind the container, shapes nW, nW+1, ...
    for c, rb in scene.get("balls", []):
        b = o.add_spheres(np.asarray(c, dtype=np.float64)[None, :], [rb], [wall_mass], m_w)
        o.L.orc_set_body_fixed(o.h, int(b), 1)
    # mesh bodies follow the container and precede the spheres: triangle shapes nW .. nW+nT-1 (one per facet)
    o.mesh_bodies = []
    if scene.get("meshes"):
        from chrono_b200 import scenes as _sc
        m_m = o.add_material(po.make_material(**(mesh_mat or mat)))
        for M in scene["meshes"]:
            q = np.asarray(M.get("rot", (1, 0, 0, 0)), dtype=np.float64)
            qc = np.concatenate([q[:1], -q[1:]])
            om_loc = _sc.quat_rotate(np.asarray(M.get("omega", (0, 0, 0)), dtype=np.float64), qc)
            b = o.add_body(M.get("mass", 1.0), (1, 1, 1), M.get("pos", (0, 0, 0)), rot=q, vel=M.get("vel", (0, 0, 0)),
                           omega=om_loc, fixed=True)
            o.add_triangles(b, m_m, M["tri"])
            o.mesh_bodies.append(b)
    first = o.add_spheres(scene["pos"], scene["radius"], sphere_mass(scene["radius"]), m_s, vel=vel, omega=omega)
    if scene.get("fixed") is not None:  # SetParticleFixed: inactive bodies
        for i in np.nonzero(scene["fixed"])[0]:
            o.L.orc_set_body_fixed(o.h, int(first + i), 1)
    o.first_sphere_body = first
    o.num_walls = len(scene["walls"])
    o.num_triangles = sum(len(M["tri"]) for M in scene.get("meshes", []))
    o.first_sphere_shape = len(scene["walls"]) + len(scene.get("balls", [])) + o.num_triangles
    return o


from chrono_b200.scenes import make_gpu  # noqa: E402,F401  (the engine-side builder lives with the product; re-exported for the tests)


def oracle_sphere_state(o):
    pos, rot, vel, om = o.state()
    f = o.first_sphere_body
    return pos[f:], vel[f:], om[f:]


def rel_err(a, b):
    """max over spheres of |a-b| / max(|b|, floor) on vector norms (not on near-zero components)."""
    a, b = np.asarray(a), np.asarray(b)
    d = np.linalg.norm(a - b, axis=-1)
    n = np.linalg.norm(b, axis=-1)
    floor = max(n.max(), 1e-300) * 1e-6
    return float((d / np.maximum(n, floor)).max())

"""Shared builders: the same scene in the CPU oracle and in the CUDA engine (through its C ABI)."""
import numpy as np

from oracle import pyoracle as po

RHO = 2000.0
MASS_COEF = 4.0 / 3.0 * np.pi * RHO


def sphere_mass(radius, mass_coef=MASS_COEF):
    """Same expression (and rounding sequence) as sphere_mass() in chrono_b200/csrc/dem_kernels.cuh."""
    r = np.asarray(radius, dtype=np.float64)
    return mass_coef * (r * r * r)


def settling_material(mu=0.4, cr=0.4, young=2e6, mu_roll=0.0, mu_spin=0.0, adhesion=0.0):
    """btest_MCORE_settling.cpp:80-92"""
    return dict(young=young, poisson=0.3, mu_s=mu, mu_roll=mu_roll, mu_spin=mu_spin, cr=cr, adhesion=adhesion)


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


def make_gpu(scene, dt=1e-3, mat=None, wall_mat=None, mesh_mat=None, wall_mass=1.0, vel=None, omega=None, gravity=(0, 0, -9.81),
             integrator=None, history_slots=16, device=0, **model):
    from chrono_b200 import dem
    mat = mat or settling_material()
    kw = dict(model)
    cfg = dem.config(device=device, dt=dt, bins=scene["bins"], gravity=gravity, mat_sphere=dem.material(**mat),
                     mat_wall=dem.material(**(wall_mat or mat)), mat_mesh=dem.material(**(mesh_mat or mat)),
                     mass_coef=MASS_COEF, wall_mass=wall_mass,
                     integrator=dem.CENTERED_DIFFERENCE if integrator is None else integrator,
                     history_slots=history_slots, **kw)
    g = dem.DemSystem(cfg)
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

"""Sphere-triangle path (ChSystemDemMesh rows a16/a17 of SURVEY 8; oracle O3 triangle_sphere) vs the CPU oracle:
a mesh is a fixed Multicore body carrying one triangle shape per facet, with the frame and velocity ApplyMeshMotion sets.

Bars: contact-pair sets (sphere-sphere, sphere-wall AND sphere-triangle) bit-exact; forces, torques, state and the
wrench on the mesh within 1e-9 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402
from test_gpu_parity import compare_step, kinematics, quat_rotate  # noqa: E402

R = 0.02


def bumpy_floor_scene(n, seed, cells_per_R=0.7, rot=None, pos=(0.0, 0.0, 0.0), vel=(0, 0, 0), omega=(0, 0, 0), mass=5.0):
    """Settling scene + a height-field mesh just above the floor wall: bottom-layer spheres overlap faces, edges and
    vertices of triangles whose edge is ~1.4 R."""
    sc = scenes.settling_scene(n, sep_factor=1.985, seed=seed)
    L = sc["box_size"][0]
    rng = np.random.default_rng(seed + 100)
    ncell = max(4, int(L / (R / cells_per_R)))
    bump = rng.uniform(0.0, 0.12 * R, size=(ncell + 1, ncell + 1))
    tri_w = scenes.heightfield_mesh(-L / 2, L / 2, -L / 2, L / 2, ncell, ncell, lambda X, Y: 0.03 * R + bump)
    rot = np.asarray(rot if rot is not None else (1.0, 0.0, 0.0, 0.0), dtype=np.float64)
    sc["meshes"] = [dict(tri=scenes.mesh_to_body_frame(tri_w, pos, rot), pos=np.asarray(pos, dtype=np.float64), rot=rot,
                         vel=np.asarray(vel, dtype=np.float64), omega=np.asarray(omega, dtype=np.float64), mass=mass)]
    return sc


def mesh_wrench_oracle(o, rot_before):
    fo, to = o.body_forces()
    b = o.mesh_bodies[0]
    return fo[b], quat_rotate(to[b:b + 1], rot_before[b:b + 1])[0]


def test_mesh_single_step_faces_edges_vertices():
    q = scenes.quat_from_axis_angle((0.3, -0.2, 1.0), 0.4)
    sc = bumpy_floor_scene(3000, 21, rot=q, pos=(0.01, -0.02, -0.005), vel=(0.05, -0.02, 0.01), omega=(0.3, -0.1, 0.2))
    vel, om = kinematics(3000, 7)
    o = common.make_oracle(sc, vel=vel, omega=om, dt=1e-4)
    _, rot_before, _, _ = o.state()
    o2, g, npairs = compare_step(sc, vel, om, steps=1, dt=1e-4, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
    ct = o2.contacts()
    nW, nT = len(sc["walls"]), o2.num_triangles
    a = (ct["shape_pair"].astype(np.uint64) >> np.uint64(32)).astype(np.int64)
    tri_contacts = (a >= nW) & (a < nW + nT)
    assert tri_contacts.sum() > 200, "scene does not exercise the mesh"
    er = ct["erad"][tri_contacts]
    assert (np.abs(er - R) < 1e-12).any() and (er < 0.9 * R).any(), "need both face and edge/vertex contacts"
    # wrench on the mesh (force, torque about the frame origin, world frame)
    fm, tm = g.mesh_wrench(0)
    f_ref, t_ref = mesh_wrench_oracle(o2, rot_before)
    assert np.linalg.norm(fm - f_ref) < 1e-9 * np.linalg.norm(f_ref)
    assert np.linalg.norm(tm - t_ref) < 1e-8 * np.linalg.norm(t_ref)
    # history rows of sphere-triangle contacts carry the triangle's shape id
    ho, hg = o2.history(), g.history()
    key_o = np.sort((ho["shape1"].astype(np.int64) << 32) | ho["shape2"].astype(np.int64))
    key_g = np.sort((hg["owner"].astype(np.int64) << 32) | hg["other"].astype(np.int64))
    assert np.array_equal(key_o, key_g)
    assert ((hg["other"] >= nW) & (hg["other"] < nW + nT)).sum() == tri_contacts.sum()


@pytest.mark.parametrize("model,mat_props,tang,roll", [
    (po.HERTZ, True, po.TANG_MULTISTEP, 0.05), (po.HOOKE, False, po.TANG_ONESTEP, 0.0), (po.FLORES, True, po.TANG_MULTISTEP, 0.0),
    (po.HERTZ, False, po.TANG_NONE, 0.0)])
def test_mesh_models_three_steps(model, mat_props, tang, roll):
    sc = bumpy_floor_scene(2000, 33, vel=(0.02, 0.0, 0.0), omega=(0.0, 0.0, 0.5))
    vel, om = kinematics(2000, 8)
    mat = common.settling_material(mu_roll=roll, mu_spin=roll / 2)
    compare_step(sc, vel, om, steps=3, dt=1e-4, mat=mat, force_model=model, use_mat_props=mat_props, tangential_mode=tang,
                 history_slots=20)


def test_moving_mesh_forty_steps():
    """ApplyMeshMotion every step (rigid rotation + translation): the mesh's travel uses up Verlet skin, lists are
    rebuilt, sphere-triangle history survives the rebuilds."""
    sc = bumpy_floor_scene(2000, 44)
    vel, om = kinematics(2000, 9, vscale=0.05, wscale=1.0)
    o = common.make_oracle(sc, vel=vel, omega=om, dt=1e-4)
    g = common.make_gpu(sc, vel=vel, omega=om, dt=1e-4)
    b = o.mesh_bodies[0]
    w = np.array([0.0, 0.0, 2.0])
    v = np.array([0.3, 0.0, 0.05])
    for it in range(40):
        t = it * 1e-4
        q = scenes.quat_from_axis_angle((0, 0, 1), 2.0 * t)
        p = v * t
        o.set_body_state(b, pos=p, rot=q, vel=v, omega=w)  # rotation about z: local omega = world omega
        g.set_mesh_motion(0, p, q, v, w)
        assert o.step(1) == 0
        g.step(1)
    pos_o, rot_o, vel_o, om_o = o.state()
    f = o.first_sphere_body
    pos_g, vel_g, om_g = g.state()
    assert common.rel_err(pos_g, pos_o[f:]) < 1e-9
    assert common.rel_err(vel_g, vel_o[f:]) < 1e-6
    st = g.stats()
    assert st["rebuilds"] >= 2, st
    ho, hg = o.history(), g.history()
    key_o = np.sort((ho["shape1"].astype(np.int64) << 32) | ho["shape2"].astype(np.int64))
    key_g = np.sort((hg["owner"].astype(np.int64) << 32) | hg["other"].astype(np.int64))
    assert np.array_equal(key_o, key_g)


def test_mesh_collision_disabled_equals_no_mesh():
    sc = bumpy_floor_scene(1500, 55)
    sc0 = dict(sc)
    sc0["meshes"] = []
    vel, om = kinematics(1500, 10)
    g = common.make_gpu(sc, vel=vel, omega=om, dt=1e-4)
    g0 = common.make_gpu(sc0, vel=vel, omega=om, dt=1e-4)
    g.enable_mesh_collision(False)
    g.step(5)
    g0.step(5)
    for a, b in zip(g.state(), g0.state()):
        assert np.array_equal(a, b)
    g.enable_mesh_collision(True)
    g.step(1)
    g0.step(1)
    assert not np.array_equal(g.state()[1], g0.state()[1])


def test_drum_mesh_encloses_spheres():
    """config[3]-style: spheres inside a closed cylinder mesh (axis y) that rotates; nothing leaves the drum and the
    GPU trajectory follows the oracle."""
    rng = np.random.default_rng(3)
    Rd, Ld = 0.25, 0.3
    tri = scenes.cylinder_drum_mesh(Rd, Ld, 48)
    pts = scenes.hcp_points((-0.12, -0.12, -0.21), (0.12, 0.12, -0.05), 2.0 * R)
    pts = pts[np.hypot(pts[:, 0], pts[:, 2]) < Rd - 1.2 * R]
    pts = pts + rng.uniform(-0.005 * R, 0.005 * R, size=pts.shape)
    n = len(pts)
    sc = dict(pos=pts, radius=np.full(n, R), walls=[], bins=(10, 10, 10), n=n,
              meshes=[dict(tri=tri, pos=np.zeros(3), rot=np.array([1.0, 0, 0, 0]), vel=np.zeros(3), omega=np.array([0, 1.0, 0]), mass=10.0)])
    o = common.make_oracle(sc, dt=1e-4)
    g = common.make_gpu(sc, dt=1e-4)
    b = o.mesh_bodies[0]
    for it in range(60):
        q = scenes.quat_from_axis_angle((0, 1, 0), 1.0 * it * 1e-4)
        o.set_body_state(b, rot=q, omega=(0, 1.0, 0))
        g.set_mesh_motion(0, None, q, None, (0, 1.0, 0))
        assert o.step(1) == 0
        g.step(1)
    pos_o, _, vel_o, _ = o.state()
    f = o.first_sphere_body
    pos_g, vel_g, _ = g.state()
    assert common.rel_err(pos_g, pos_o[f:]) < 1e-9
    assert common.rel_err(vel_g, vel_o[f:]) < 1e-6
    assert (np.hypot(pos_g[:, 0], pos_g[:, 2]) < Rd).all()

"""Parity holes named by the round-1 review: wall velocity entering the force, plane and z-cylinder walls against the
closed-form Hertz force, the FORWARD_EULER integrator against its closed form.

* moving walls: every wall of the box gets a velocity; the oracle's container body (body 0) gets the same through
  set_body_state -- the relative velocity at a wall contact changes both the damping and the friction force
  (ChIterativeSolverMulticoreSMC.cpp:141-156).  Bars as everywhere: pair sets bit-exact, forces / state 1e-9.
* plane, z-cylinder (inside and outside): one sphere, one step, no gravity: |dv| = F_hertz / m * h, direction = wall normal.
* FORWARD_EULER (ChDemSMC.cuh integrateSpheres, CHDEM_TIME_INTEGRATOR::FORWARD_EULER): x+ = x + h v, v+ = v + h a."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402
from test_gpu_parity import kinematics, quat_rotate  # noqa: E402
from test_gpu_bc import one_sphere, hertz_speed, R  # noqa: E402


@pytest.mark.parametrize("tang", [po.TANG_MULTISTEP, po.TANG_ONESTEP])
def test_moving_walls_match_the_oracle(tang):
    n = 3000
    scene = scenes.settling_scene(n, sep_factor=1.985, seed=61)
    # the lattice starts 0.01 R clear of the container: push every wall 0.03 R inwards so that the outer spheres touch it
    centre = np.array([0.0, 0.0, scene["box_size"][2] / 2])
    walls = []
    for p, h in scene["walls"]:
        a = int(np.argmin(h))
        p = np.array(p, dtype=np.float64)
        p[a] -= np.sign(p[a] - centre[a]) * 0.03 * R
        walls.append((p, h))
    scene["walls"] = walls
    vel, om = kinematics(n, 17)
    wall_v = np.array([0.35, -0.2, 0.15])
    kw = dict(dt=1e-4, force_model=po.HERTZ, tangential_mode=tang)
    o = common.make_oracle(scene, vel=vel, omega=om, **kw)
    g = common.make_gpu(scene, vel=vel, omega=om, **kw)
    o.set_body_state(0, vel=wall_v)  # the container: a fixed body that carries a velocity
    from chrono_b200 import dem
    for w in range(g.num_walls):
        g._ck(g.L.dem_b200_set_wall_state(g.h, w, None, wall_v.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double))))
    g.enable_recording(True, max_pairs=40 * n)
    f = o.first_sphere_body
    nwall_contacts = 0
    for it in range(3):
        _, rot_before, _, _ = o.state()
        assert o.step(1) == 0
        ct = o.contacts()
        fo, to = o.body_forces()
        g.step(1)
        assert np.array_equal(np.sort(ct["shape_pair"].astype(np.uint64)), np.sort(g.pairs()))
        nwall_contacts += int(((ct["shape_pair"].astype(np.uint64) >> np.uint64(32)) < len(scene["walls"])).sum())
        fg, tg = g.forces()
        assert common.rel_err(fg, fo[f:]) < 1e-9, ("force", it, common.rel_err(fg, fo[f:]))
        assert common.rel_err(tg, quat_rotate(to[f:], rot_before[f:])) < 1e-8, ("torque", it)
        pos_o, rot_o, vel_o, om_o = o.state()
        pos_g, vel_g, om_g = g.state()
        assert common.rel_err(pos_g, pos_o[f:]) < 1e-9 and common.rel_err(vel_g, vel_o[f:]) < 1e-9 * (1 + 10 * it)
    assert nwall_contacts > 300, "the scene must put spheres on the walls"
    # and the wall velocity matters: the same step with walls at rest gives different forces on the wall spheres
    g0 = common.make_gpu(scene, vel=vel, omega=om, **kw)
    g0.enable_recording(True, max_pairs=40 * n)
    g0.step(1)
    g1 = common.make_gpu(scene, vel=vel, omega=om, **kw)
    import ctypes
    for w in range(g1.num_walls):
        g1._ck(g1.L.dem_b200_set_wall_state(g1.h, w, None, wall_v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    g1.enable_recording(True, max_pairs=40 * n)
    g1.step(1)
    assert np.abs(g1.forces()[0] - g0.forces()[0]).max() > 1e-3


def test_plane_wall_hertz_force():
    delta, dt = 0.035 * R, 1e-5
    nrm = np.array([0.2, -0.3, 0.9])
    nrm /= np.linalg.norm(nrm)
    p0 = np.array([0.1, 0.2, -0.05])
    pos = p0 + nrm * (R - delta) + np.cross(nrm, [1.0, 0, 0]) * 0.37  # anywhere on the plane
    v = one_sphere(lambda g: g.add_plane_wall(p0, nrm), pos, dt)
    expect = hertz_speed(delta, R, dt)  # face contact: effective radius = r (ChNarrowphasePRIMS box face convention)
    assert abs(np.linalg.norm(v) - expect) < 1e-9 * expect
    assert np.linalg.norm(v / np.linalg.norm(v) - nrm) < 1e-9
    # a sphere that does not reach the plane feels nothing
    assert np.linalg.norm(one_sphere(lambda g: g.add_plane_wall(p0, nrm), p0 + nrm * (R * 1.001), dt)) == 0.0


@pytest.mark.parametrize("inside", [True, False])
def test_zcylinder_wall_hertz_force(inside):
    import ctypes as C
    Rc, delta, dt = 0.6, 0.03 * R, 1e-5
    phi = 1.1
    axis = np.array([0.05, -0.1, 0.0])
    rho = (Rc - R + delta) if inside else (Rc + R - delta)
    pos = axis + np.array([rho * math.cos(phi), rho * math.sin(phi), 0.3])

    def build(g):
        c = np.ascontiguousarray(axis)
        g._ck(g.L.dem_b200_add_zcylinder_wall(g.h, c.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(Rc), int(inside)))

    v = one_sphere(build, pos, dt)
    expect = hertz_speed(delta, R, dt)
    radial = np.array([math.cos(phi), math.sin(phi), 0.0]) * (-1.0 if inside else 1.0)  # pushed away from the wall
    assert abs(np.linalg.norm(v) - expect) < 1e-9 * expect
    assert np.linalg.norm(v / np.linalg.norm(v) - radial) < 1e-9


def test_forward_euler_closed_form():
    """Free flight under gravity with FORWARD_EULER: x_k = x0 + k h v0 + g h^2 k (k-1) / 2, v_k = v0 + k h g (positions lag the
    velocity update by one step); then one step in contact: x+ uses the OLD velocity, v+ the contact acceleration."""
    from chrono_b200 import dem
    h, k = 1e-3, 25
    pos = np.array([[0.0, 0.0, 1.0], [0.5, 0.1, 1.3]])
    v0 = np.array([[0.3, -0.2, 0.1], [0.0, 0.0, -1.0]])
    gvec = np.array([0.0, 0.0, -9.81])
    scene = dict(pos=pos, radius=np.full(2, R), walls=scenes.box_container((4, 4, 3), 0.2, (0, 0, 1.5)), bins=(10, 10, 8), n=2)
    g = common.make_gpu(scene, vel=v0, dt=h, integrator=dem.FORWARD_EULER)
    g.step(k)
    p, v, w = g.state()
    assert np.allclose(v, v0 + k * h * gvec, rtol=0, atol=1e-13)
    assert np.allclose(p, pos + k * h * v0 + gvec * h * h * k * (k - 1) / 2, rtol=0, atol=1e-12)
    # one step in contact with the floor (Hertz, no damping contribution at zero normal speed, no gravity)
    delta = 0.02 * R
    scene = dict(pos=np.array([[0.0, 0.0, R - delta]]), radius=np.full(1, R), walls=scenes.box_container((4, 4, 3), 0.2, (0, 0, 1.5)),
                 bins=(10, 10, 8), n=1)
    vin = np.array([[0.2, 0.0, 0.0]])
    g = common.make_gpu(scene, vel=vin, dt=1e-5, gravity=(0, 0, 0), integrator=dem.FORWARD_EULER, tangential_mode=dem.TANG_NONE,
                        mat=common.settling_material(mu=0.0))
    g.step(1)
    p, v, w = g.state()
    assert np.allclose(p[0], [0.2 * 1e-5, 0.0, R - delta], rtol=0, atol=1e-15)  # the position moved with the old velocity only
    assert abs(v[0, 2] - hertz_speed(delta, R, 1e-5)) < 1e-9 * hertz_speed(delta, R, 1e-5) and v[0, 0] == 0.2

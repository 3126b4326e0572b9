"""Boundary conditions beyond the box container: ball, cone, plane, z-cylinder.

* ball obstacle == Multicore's sphere_sphere against a fixed sphere body: full parity harness (bit-exact pair sets).
* cone and cavity have no Multicore counterpart: single-step checks against the analytic Hertz force / an equivalent plane."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402
from test_gpu_parity import compare_step, kinematics  # noqa: E402

R = 0.02


def test_ball_obstacle_matches_fixed_sphere_body():
    sc = scenes.settling_scene(2500, sep_factor=1.985, seed=31)
    top = (sc["pos"][:, 2] + sc["radius"]).max()
    Rb = 4.0 * R
    sc["balls"] = [((0.003, -0.002, top + Rb - 0.08 * R), Rb), ((0.15, 0.1, 0.6 * R + Rb * 0.3), Rb * 0.5)]
    vel, om = kinematics(2500, 12)
    o, g, npairs = compare_step(sc, vel, om, steps=3, dt=1e-4, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
    ct = o.contacts()
    nW = len(sc["walls"])
    a = (ct["shape_pair"].astype(np.uint64) >> np.uint64(32)).astype(np.int64)
    assert ((a >= nW) & (a < nW + 2)).sum() >= 3, "no sphere touches a ball boundary"


def one_sphere(build, pos, dt=1e-5, young=2e6):
    from chrono_b200 import dem
    mat = common.settling_material(young=young)
    cfg = dem.config(dt=dt, gravity=(0, 0, 0), mat_sphere=dem.material(**mat), mat_wall=dem.material(**mat),
                     mass_coef=common.MASS_COEF, wall_mass=1e30, force_model=dem.HERTZ, tangential_mode=dem.TANG_NONE)
    g = dem.DemSystem(cfg)
    build(g)
    g.set_spheres(np.asarray([pos], dtype=np.float64), [R])
    g.initialize()
    g.step(1)
    return g.state()[1][0]


def hertz_speed(delta, erad, dt, young=2e6, nu=0.3):
    f32 = np.float32
    inv_E = (f32(1) - f32(nu) * f32(nu)) / f32(young) + (f32(1) - f32(nu) * f32(nu)) / f32(young)
    E = float(f32(1) / inv_E)
    kn = (2.0 / 3.0) * 2.0 * E * math.sqrt(erad * delta)
    m = common.sphere_mass(R)
    return kn * delta / m * dt


def test_cavity_hertz_force():
    Rb, delta, dt = 0.5, 0.03 * R, 1e-5
    d = np.array([0.3, -0.4, -math.sqrt(1 - 0.25)])
    pos = d * (Rb - R + delta)
    v = one_sphere(lambda g: g.add_sphere_wall((0, 0, 0), Rb, spheres_outside=False), pos, dt)
    expect = hertz_speed(delta, Rb * R / (Rb - R), dt)
    assert abs(np.linalg.norm(v) - expect) < 1e-9 * expect
    assert np.linalg.norm(v / np.linalg.norm(v) + d) < 1e-9  # pushed towards the centre


def test_cone_equals_tangent_plane():
    slope, delta, dt = 0.75, 0.04 * R, 1e-5
    phi = 0.7
    rho = 0.3
    p_surf = np.array([rho * math.cos(phi), rho * math.sin(phi), slope * rho])
    n = np.array([-slope * math.cos(phi), -slope * math.sin(phi), 1.0])
    n /= np.linalg.norm(n)  # inward (above the surface)
    pos = p_surf + n * (R - delta)
    v_cone = one_sphere(lambda g: g.add_zcone_wall((0, 0, 0), slope, -1.0, 1.0, spheres_above=True), pos, dt)
    v_plane = one_sphere(lambda g: g.add_plane_wall(p_surf, n), pos, dt)
    assert np.linalg.norm(v_cone) > 0
    assert np.linalg.norm(v_cone - v_plane) < 1e-9 * np.linalg.norm(v_plane)
    # outside its height range, or on the wrong side, the cone exerts nothing
    assert np.linalg.norm(one_sphere(lambda g: g.add_zcone_wall((0, 0, 0), slope, -1.0, 0.1, spheres_above=True), pos, dt)) == 0.0
    assert np.linalg.norm(one_sphere(lambda g: g.add_zcone_wall((0, 0, 0), slope, -1.0, 1.0, spheres_above=False), pos, dt)) == 0.0


def test_hopper_holds_a_bed():
    """A few hundred spheres poured into a cone with a plane floor under the tip come to rest inside it."""
    from chrono_b200 import dem
    rng = np.random.default_rng(2)
    pts = scenes.hcp_points((-0.2, -0.2, 0.45), (0.2, 0.2, 0.65), 2.3 * R)
    pts = pts[np.hypot(pts[:, 0], pts[:, 1]) < 0.2]
    pts = pts + rng.uniform(-0.01 * R, 0.01 * R, size=pts.shape)
    mat = common.settling_material(mu=0.5, cr=0.2)
    cfg = dem.config(dt=1e-4, gravity=(0, 0, -9.81), mat_sphere=dem.material(**mat), mat_wall=dem.material(**mat),
                     mass_coef=common.MASS_COEF, wall_mass=1e30, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP,
                     history_slots=20)
    g = dem.DemSystem(cfg)
    g.add_zcone_wall((0, 0, 0), 1.0, 0.05, 2.0, spheres_above=True)
    g.add_plane_wall((0, 0, 0.05), (0, 0, 1))
    g.set_spheres(pts, np.full(len(pts), R))
    g.initialize()
    g.step(12000)
    p, v, _ = g.state()
    assert np.quantile(np.linalg.norm(v, axis=1), 0.99) < 0.05
    rho = np.hypot(p[:, 0], p[:, 1])
    assert (p[:, 2] > 0.05 + 0.9 * R).all() and (p[:, 2] - rho > -0.05 * R * math.sqrt(2)).all()  # above floor and cone
    assert p[:, 2].max() < 0.45  # the cloud has fallen into the hopper


def test_fixed_spheres_are_inactive_bodies():
    """SetParticleFixed (terrain made of fixed particles, demo_DEM_fixedTerrain): a fixed sphere is an inactive Multicore body --
    no contact between two fixed spheres or between a fixed sphere and a wall, it pushes but is not pushed."""
    sc = scenes.settling_scene(3000, sep_factor=1.985, seed=41)
    z = sc["pos"][:, 2]
    sc["fixed"] = (z < np.quantile(z, 0.3)).astype(np.uint8)  # the bottom layers are terrain
    vel, om = kinematics(3000, 13)
    vel[sc["fixed"] != 0] = 0.0
    om[sc["fixed"] != 0] = 0.0
    o, g, npairs = compare_step(sc, vel, om, steps=3, dt=1e-4, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
    p, v, _ = g.state()
    assert np.array_equal(p[sc["fixed"] != 0], sc["pos"][sc["fixed"] != 0]) and not v[sc["fixed"] != 0].any()

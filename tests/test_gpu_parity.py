"""CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): bin assignment and contact-pair lists bit-exact; single-step forces and
positions within 1e-9 relative (fp64, evaluated on vector norms); multi-step trajectories stay within a much
looser bound because the two implementations sum in different orders."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402

TOL = 1e-9


def quat_rotate(v, q):
    w, u = q[:, :1], q[:, 1:]
    t = 2 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def kinematics(n, seed, vscale=0.2, wscale=5.0):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n, 3)) * vscale, rng.normal(size=(n, 3)) * wscale


def compare_step(scene, vel, omega, steps=1, tol=TOL, **kw):
    from chrono_b200 import dem
    o = common.make_oracle(scene, vel=vel, omega=omega, **kw)
    g = common.make_gpu(scene, vel=vel, omega=omega, **kw)
    nW, n = o.first_sphere_shape, scene["n"]  # first sphere shape (walls, ball boundaries, mesh triangles, spheres)
    g.enable_recording(True, max_pairs=40 * n)
    for it in range(steps):
        # one step on both sides; the oracle keeps the bins / contacts / forces it used during that step
        # (orc_eval is NOT called separately: with MultiStep history an extra evaluation would advance the history)
        _, rot_before, _, _ = o.state()
        assert o.step(1) == 0
        gmin_o, gmax_o = o.shape_bins()
        ct = o.contacts()
        fo, to = o.body_forces()
        g.step(1)
        # ---- bin assignment: bit-exact
        gmin_g, gmax_g = g.bins()
        assert np.array_equal(gmin_g, gmin_o[nW:]), "HashMin bins differ at step %d" % it
        assert np.array_equal(gmax_g, gmax_o[nW:]), "HashMax bins differ at step %d" % it
        og, ob, oib, _ = o.grid()
        gg, gb, gib = g.grid()
        assert np.array_equal(og, gg) and np.array_equal(ob, gb) and np.array_equal(oib, gib)
        # ---- contact pair list: bit-exact as a set of (shapeA<<32 | shapeB)
        po_pairs = np.sort(ct["shape_pair"].astype(np.uint64))
        pg_pairs = np.sort(g.pairs())
        assert len(po_pairs) == len(pg_pairs), (len(po_pairs), len(pg_pairs))
        assert np.array_equal(po_pairs, pg_pairs), "contact pair sets differ at step %d" % it
        # ---- per-sphere contact force / torque
        fg, tg = g.forces()
        f_o = fo[o.first_sphere_body:]
        t_o = quat_rotate(to[o.first_sphere_body:], rot_before[o.first_sphere_body:])  # local -> global
        assert common.rel_err(fg, f_o) < tol, ("force", it, common.rel_err(fg, f_o))
        assert common.rel_err(tg, t_o) < tol * 10, ("torque", it, common.rel_err(tg, t_o))
        # ---- state after the step
        pos_o, rot_o, vel_o, om_o = o.state()
        f = o.first_sphere_body
        pos_g, vel_g, om_g = g.state()
        assert common.rel_err(pos_g, pos_o[f:]) < tol, ("pos", it)
        assert common.rel_err(vel_g, vel_o[f:]) < tol * (1 + 10 * it), ("vel", it, common.rel_err(vel_g, vel_o[f:]))
        assert common.rel_err(om_g, quat_rotate(om_o[f:], rot_o[f:])) < tol * (10 + 10 * it), ("omega", it)
    return o, g, len(po_pairs)


def test_single_step_10k_hertz_multistep():
    """config[0]-sized case: 10k monodisperse spheres, Hertz + MultiStep history, five box walls."""
    scene = scenes.settling_scene(10000, sep_factor=1.98, seed=12346)
    vel, om = kinematics(10000, 1)
    o, g, npairs = compare_step(scene, vel, om, steps=1, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
    assert npairs > 25000  # ~6 contacts per sphere
    # history written by the step: same (owner, other) keys and displacement within 1e-9
    ho, hg = o.history(), g.history()
    nW = len(scene["walls"])
    key_o = (ho["shape1"].astype(np.int64) << 32) | ho["shape2"].astype(np.int64)
    key_g = (hg["owner"].astype(np.int64) << 32) | hg["other"].astype(np.int64)
    io, ig = np.argsort(key_o), np.argsort(key_g)
    assert np.array_equal(key_o[io], key_g[ig])
    assert common.rel_err(hg["disp"][ig], ho["disp"][io]) < 1e-8


@pytest.mark.parametrize("model,mat_props,tang", [
    (po.HERTZ, True, po.TANG_ONESTEP), (po.HERTZ, False, po.TANG_MULTISTEP), (po.HOOKE, True, po.TANG_MULTISTEP),
    (po.HOOKE, False, po.TANG_NONE), (po.FLORES, True, po.TANG_MULTISTEP), (po.PLAINCOULOMB, True, po.TANG_NONE)])
def test_force_models_three_steps(model, mat_props, tang):
    scene = scenes.settling_scene(3000, sep_factor=1.985, seed=5)
    vel, om = kinematics(3000, 2)
    compare_step(scene, vel, om, steps=3, dt=1e-4, force_model=model, use_mat_props=mat_props, tangential_mode=tang)


def test_polydisperse_rolling_spinning_adhesion():
    """config[2]-style physics: radii U(0.8,1.2)R, rolling + spinning resistance, constant adhesion."""
    scene = scenes.settling_scene(4000, polydisperse=(0.8, 1.2), sep_factor=1.6, seed=9)
    vel, om = kinematics(4000, 3)
    mat = common.settling_material(mu_roll=0.05, mu_spin=0.02, adhesion=1e-3)
    compare_step(scene, vel, om, steps=3, dt=1e-4, mat=mat, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP,
                 history_slots=20)


def test_fifty_steps_history_persistence():
    """History entries must survive re-sorting: 50 steps, compare trajectories and the surviving history keys."""
    scene = scenes.settling_scene(2000, sep_factor=1.99, seed=11)
    vel, om = kinematics(2000, 4, vscale=0.05, wscale=1.0)
    o = common.make_oracle(scene, vel=vel, omega=om, dt=1e-4)
    g = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4)
    assert o.step(50) == 0
    g.step(50)
    pos_o, rot_o, vel_o, om_o = o.state()
    f = o.first_sphere_body
    pos_g, vel_g, om_g = g.state()
    assert common.rel_err(pos_g, pos_o[f:]) < 1e-9
    assert common.rel_err(vel_g, vel_o[f:]) < 1e-6
    ho, hg = o.history(), g.history()
    key_o = np.sort((ho["shape1"].astype(np.int64) << 32) | ho["shape2"].astype(np.int64))
    key_g = np.sort((hg["owner"].astype(np.int64) << 32) | hg["other"].astype(np.int64))
    assert np.array_equal(key_o, key_g)
    io = np.argsort((ho["shape1"].astype(np.int64) << 32) | ho["shape2"].astype(np.int64))
    ig = np.argsort((hg["owner"].astype(np.int64) << 32) | hg["other"].astype(np.int64))
    assert np.allclose(hg["duration"][ig], ho["duration"][io], rtol=0, atol=1e-12)


def test_reference_fixture_pairs():
    """Contact pairs of the committed reference-generated scene (tests/golden/ref_vectors.npz), no oracle involved."""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
    scene = scenes.settling_scene(int(G["scene_n"][0]), sep_factor=1.98, seed=int(G["scene_seed"][0]))
    g = common.make_gpu(scene)
    g.enable_recording(True, max_pairs=40 * scene["n"])
    g.step(1)
    assert np.array_equal(np.sort(g.pairs()), np.sort(G["scene_ct_shape"].astype(np.uint64)))
    gg, gb, gib = g.grid()
    assert np.array_equal(gg, G["scene_origin"]) and np.array_equal(gib, G["scene_inv_bin_size"])


def test_determinism_and_graph_path():
    """Two runs give bit-identical states; the CUDA-graph path (even step counts) equals direct launches."""
    scene = scenes.settling_scene(5000, sep_factor=1.99, seed=21)
    vel, om = kinematics(5000, 6)
    a = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4)
    b = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4)
    a.step(20)             # graph replays
    for _ in range(20):    # direct launches
        b.step(1)
    for x, y in zip(a.state(), b.state()):
        assert np.array_equal(x, y)


def test_ragged_and_tiny_inputs():
    """1 sphere, 2 spheres, 33 spheres (not a multiple of any block size); free fall matches the closed form."""
    from chrono_b200 import dem
    for n in (1, 2, 33):
        pos = np.array([[0.1 * i, 0.0, 1.0] for i in range(n)])
        scene = dict(pos=pos, radius=np.full(n, 0.02), walls=scenes.box_container((4, 4, 2), 0.2, (0, 0, 1.0)),
                     bins=(10, 10, 5), n=n)
        g = common.make_gpu(scene, dt=1e-3)
        g.step(10)
        p, v, w = g.state()
        # semi-implicit Euler: v_k = -g k h ; z_k = z0 - g h^2 k(k+1)/2
        assert np.allclose(v[:, 2], -9.81 * 10 * 1e-3, rtol=1e-13)
        assert np.allclose(p[:, 2], 1.0 - 9.81 * 1e-6 * 55, rtol=1e-13)


def test_reductions_and_single_sphere_access():
    from chrono_b200 import dem
    scene = scenes.settling_scene(1000, seed=3)
    vel, om = kinematics(1000, 8)
    g = common.make_gpu(scene, vel=vel, omega=om)
    pos, v, w = g.state()
    assert g.reduce(dem.RED_MAX_Z) == pos[:, 2].max()
    assert g.reduce(dem.RED_MIN_Z) == pos[:, 2].min()
    m = common.sphere_mass(scene["radius"])
    ke = 0.5 * (m * (v ** 2).sum(1)).sum() + 0.5 * (0.4 * m * scene["radius"] ** 2 * (w ** 2).sum(1)).sum()
    assert abs(g.reduce(dem.RED_KE) - ke) < 1e-12 * ke
    assert g.reduce(dem.RED_COUNT_ABOVE_Z, 0.05) == (pos[:, 2] > 0.05).sum()
    p7, v7, w7 = g.sphere(7)
    assert np.array_equal(p7, pos[7]) and np.array_equal(v7, v[7]) and np.array_equal(w7, w[7])


@pytest.mark.parametrize("integ", ["CENTERED_DIFFERENCE", "EXTENDED_TAYLOR", "CHUNG"])
def test_acceleration_of_the_last_step(integ):
    """dem_b200_get_accel (GetParticleLinAcc, the fx,fy,fz columns of WriteParticleFile: ChSystemDem_impl.cpp:1290-1296,
    322-327) against g + F / m with F the recorded contact force of the same step; zero before the first step and for
    fixed spheres; graph path == recorded path."""
    from chrono_b200 import dem
    n = 3000
    scene = scenes.settling_scene(n, sep_factor=1.985, seed=5)
    fixed = np.zeros(n, dtype=np.uint8)
    fixed[::17] = 1
    scene["fixed"] = fixed
    vel, om = kinematics(n, 21)
    g = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4, integrator=getattr(dem, integ))
    assert not g.accel().any()
    g.enable_recording(True, max_pairs=40 * n)
    g.step(3)
    F, _ = g.forces()
    m = np.broadcast_to(common.sphere_mass(scene["radius"]), (n,))[:, None]
    want = np.array([0, 0, -9.81]) + F / m
    want[fixed != 0] = 0
    got = g.accel()
    assert np.abs(F).max() > 0
    assert common.rel_err(got, want) < 1e-8
    assert not got[fixed != 0].any()
    # the same three steps through the step graph
    g2 = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4, integrator=getattr(dem, integ))
    g2.step(3)
    assert common.rel_err(g2.accel(), got) < 1e-8
    # replacing the state makes the last step's acceleration meaningless: zero again
    p, v, w = g2.state()
    g2.set_state(vel=v)
    assert not g2.accel().any()


def test_bins_smaller_than_a_sphere():
    """Multicore registers a shape in every bin its AABB overlaps, so a resolution with bins smaller than a sphere
    diameter is legal there.  The engine's own search grid is independent of bins_per_axis: pair lists, bin ranges
    and forces must still match the oracle."""
    scene = scenes.settling_scene(600, sep_factor=1.97, seed=3)
    scene = dict(scene, bins=(60, 60, 30))
    vel, om = kinematics(600, 12)
    compare_step(scene, vel, om, steps=2, dt=1e-4)


@pytest.mark.parametrize("skin", [0.0, 0.1, 2.0])
def test_verlet_skin_does_not_change_a_bit(skin):
    """The candidate lists only bound the search: any skin (0 = rebuild every step) gives bit-identical states, and
    the rebuild count drops as the skin grows."""
    from chrono_b200 import dem
    scene = scenes.settling_scene(3000, sep_factor=1.99, seed=31)
    vel, om = kinematics(3000, 7, vscale=0.3)
    ref = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4, verlet_skin=0.25 * 0.02)
    g = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4, verlet_skin=skin * 0.02, neighbor_slots=64)
    ref.step(60)
    g.step(60)
    for x, y in zip(ref.state(), g.state()):
        assert np.array_equal(x, y)
    st = g.stats()
    assert st["steps"] == 60
    if skin == 0.0:
        assert st["rebuilds"] == 60
    else:
        assert 1 <= st["rebuilds"] < 60
    hr, hg = ref.history(), g.history()
    kr = np.argsort((hr["owner"].astype(np.int64) << 32) | hr["other"]); kg = np.argsort((hg["owner"].astype(np.int64) << 32) | hg["other"])
    assert np.array_equal(hr["owner"][kr], hg["owner"][kg]) and np.array_equal(hr["disp"][kr], hg["disp"][kg])


def test_host_round_trip_costs_no_rebuild_and_no_bit():
    """dem_b200_set_state / dem_b200_advance_host: positions handed back unchanged use up no Verlet skin (no rebuild: the
    co-simulation round trip of the e2e bench figure); positions that jumped count like a step's displacement, so the lists
    are only rebuilt once the jumps no longer fit into the skin; dem_b200_request_rebuild forces a rebuild.  None of it
    changes a bit of the trajectory."""
    n = 3000
    scene = scenes.settling_scene(n, sep_factor=1.99, seed=13)
    vel, om = kinematics(n, 9, vscale=0.2)
    ref = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4)
    ref.step(40)
    g = common.make_gpu(scene, vel=vel, omega=om, dt=1e-4)
    g.step(10)
    r0 = g.stats()["rebuilds"]
    p, v, w = g.state()
    g.set_state(pos=p, vel=v, omega=w)
    g.step(1)
    assert g.stats()["rebuilds"] == r0, "unchanged positions must not rebuild"
    po, vo, wo = np.empty_like(p), np.empty_like(p), np.empty_like(p)
    p, v, w = g.state()
    g.advance_host(p, v, w, 9, po, vo, wo)  # steps 12 .. 20 through the host-buffer call
    g.request_rebuild()
    r1 = g.stats()["rebuilds"]
    g.step(1)
    assert g.stats()["rebuilds"] == r1 + 1
    g.step(19)
    for x, y in zip(ref.state(), g.state()):
        assert np.array_equal(x, y)
    # both engines are in the same state now (history included).  Small jumps of every sphere: one engine keeps its lists,
    # the other is told to rebuild; five steps later they still agree bit for bit
    ref.request_rebuild()
    g.request_rebuild()
    ref.step(1)
    g.step(1)  # travel starts from zero in both
    p, v, w = g.state()
    p2 = p + np.random.default_rng(3).normal(size=p.shape) * 1e-5
    ra, rb = g.stats()["rebuilds"], ref.stats()["rebuilds"]
    g.set_state(pos=p2)
    ref.set_state(pos=p2)
    ref.request_rebuild()
    g.step(1)
    ref.step(1)
    assert g.stats()["rebuilds"] == ra and ref.stats()["rebuilds"] == rb + 1
    g.step(4)
    ref.step(4)
    for x, y in zip(ref.state(), g.state()):
        assert np.array_equal(x, y)
    # a jump far beyond the skin: the next step rebuilds on its own
    p, v, w = g.state()
    p[5, 2] += 3.0
    ra = g.stats()["rebuilds"]
    g.set_state(pos=p)
    g.step(1)
    assert g.stats()["rebuilds"] == ra + 1


def test_history_overflow_is_reported():
    """More simultaneous contacts than history_slots must fail loudly, not silently drop a contact."""
    from chrono_b200 import dem
    scene = scenes.settling_scene(500, sep_factor=1.9, seed=3)
    g = common.make_gpu(scene, history_slots=4)
    with pytest.raises(dem.DemError) as e:
        g.step(1)
    assert e.value.code == -4


def test_slab_decomposition_two_gpus_bit_identical():
    """2-rank slab run (NCCL halo exchange, migration with history) == single-GPU run, bit for bit (tests/mgpu_check.py)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(here, "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

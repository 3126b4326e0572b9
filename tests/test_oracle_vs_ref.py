"""The oracle restatement against the REFERENCE's own object code (oracle/_ref/libchrono_ref.so, compiled from
/root/reference by oracle/Makefile).  Bit-exact comparisons; skipped where the reference is not mounted
(e.g. on the GPU box), where tests/golden/ref_vectors.npz takes over (test_oracle_fixtures.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from chrono_b200 import scenes
import dem_common as common

pytestmark = pytest.mark.skipif(not po.ref_available(), reason="reference objects not built (no /root/reference)")


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def same(a, b):
    if a is None or b is None:
        assert a is None and b is None
        return
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (k, a[k], b[k])


def test_rotate_bitexact():
    rng = np.random.default_rng(1)
    R = po.ref()
    for _ in range(500):
        v, q = rng.normal(size=3), rand_quat(rng)
        for x, y in zip(po.rotate(v, q), R.rotate(v, q)):
            assert np.array_equal(x, y)


def test_prims_bitexact_random():
    rng = np.random.default_rng(2)
    O, R = po.orc_prims(), po.ref().prims
    hits = 0
    for _ in range(3000):
        p1, p2 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        r1, r2 = rng.uniform(0.2, 0.9, 2)
        sep = rng.choice([0.0, 0.05])
        a, b = O.sphere_sphere(p1, r1, p2, r2, sep), R.sphere_sphere(p1, r1, p2, r2, sep)
        same(a, b)
        hits += a is not None
        q = rand_quat(rng)
        hd = rng.uniform(0.2, 1.0, 3)
        a, b = O.box_sphere(p1, q, hd, p2 * 2, r2, sep), R.box_sphere(p1, q, hd, p2 * 2, r2, sep)
        same(a, b)
        A, B, Cc = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        a, b = O.triangle_sphere(A, B, Cc, p2, r2, sep), R.triangle_sphere(A, B, Cc, p2, r2, sep)
        same(a, b)
        ea, ra = O.snap_to_triangle(A, B, Cc, p2)
        eb, rb = R.snap_to_triangle(A, B, Cc, p2)
        assert ea == eb and np.array_equal(ra, rb)
        ca, la = O.snap_to_box(hd, p2 * 1.5)
        cb, lb = R.snap_to_box(hd, p2 * 1.5)
        assert ca == cb and np.array_equal(la, lb)
    assert hits > 100


def _scene_arrays(o, scene):
    """Describe the oracle's scene to the reference shim."""
    nW = len(scene["walls"])
    n = scene["n"]
    ns = nW + n
    types = np.array([po.SHAPE_BOX] * nW + [po.SHAPE_SPHERE] * n, dtype=np.int32)
    bodies = np.array([0] * nW + list(range(1, n + 1)), dtype=np.int32)
    lpos = np.zeros((ns, 3))
    dims = np.zeros((ns, 3))
    for k, (p, h) in enumerate(scene["walls"]):
        lpos[k], dims[k] = p, h
    dims[nW:, 0] = scene["radius"]
    lrot = np.tile([1.0, 0, 0, 0], (ns, 1))
    tri = np.zeros((ns, 9))
    pos, rot, _, _ = o.state()
    nb = len(pos)
    active = np.ones(nb, dtype=np.int8)
    active[0] = 0
    collide = np.ones(nb, dtype=np.int8)
    return types, bodies, lpos, lrot, dims, tri, pos, rot, active, collide


@pytest.mark.parametrize("n,poly", [(400, None), (1500, (0.8, 1.2))])
def test_collision_pipeline_bitexact(n, poly):
    """AABB -> ChBroadphase::Process -> ChNarrowphase::Process of the reference vs the restatement: grid, bin CSR,
    candidate pairs and contacts all bit-identical (order included, the reference sort being stable here)."""
    scene = scenes.settling_scene(n, polydisperse=poly, sep_factor=1.6 if poly else 1.98)  # slightly overlapping lattice -> contacts
    o = common.make_oracle(scene)
    mn, mx = o.generate_aabb()
    o.eval()
    types, bodies, lpos, lrot, dims, tri, pos, rot, active, collide = _scene_arrays(o, scene)
    r = po.ref().collision(types, bodies, lpos, lrot, dims, tri, pos, rot, active, collide, mn, mx, scene["bins"])
    og, ob, oib, _ = o.grid()
    assert np.array_equal(og, r["origin"]) and np.array_equal(ob, r["bin_size"]) and np.array_equal(oib, r["inv_bin_size"])
    act, start, num = o.bin_csr()
    assert np.array_equal(act, r["bin_active"])
    assert np.array_equal(start, r["bin_start_index"])
    assert np.array_equal(num, r["bin_aabb_number"])
    assert np.array_equal(o.pairs(), r["pairs"])
    oc, rc = o.contacts(), r["contacts"]
    assert len(oc["depth"]) > n // 2  # a real contact network
    for k in oc:
        assert np.array_equal(oc[k], rc[k]), k


@pytest.mark.parametrize("model", [po.HERTZ, po.HOOKE, po.FLORES, po.PLAINCOULOMB])
@pytest.mark.parametrize("mat_props", [True, False])
@pytest.mark.parametrize("tang", [po.TANG_NONE, po.TANG_ONESTEP, po.TANG_MULTISTEP])
def test_contact_force_bitexact(model, mat_props, tang):
    """function_CalcContactForces (reference object code) vs the restatement on random contacts, including
    rolling / spinning friction, the three adhesion laws, new and persistent MultiStep history, both orientations."""
    rng = np.random.default_rng(100 + model * 10 + tang)
    for it in range(300):
        adh = int(rng.integers(0, 3))
        s = po.make_settings(force_model=model, adhesion_model=adh, tangential_mode=tang, use_mat_props=mat_props,
                             dt=float(rng.choice([1e-3, 1e-4])))
        m1 = po.make_material(young=float(rng.uniform(1e5, 1e7)), poisson=0.3, mu_s=float(rng.uniform(0, 0.8)),
                              mu_roll=float(rng.choice([0.0, 0.05])), mu_spin=float(rng.choice([0.0, 0.02])),
                              cr=float(rng.uniform(0.05, 0.95)), adhesion=float(rng.choice([0.0, 0.3])),
                              adhesion_dmt=0.1, adhesion_perko=0.2, kn=2e5, kt=1e5, gn=40, gt=20)
        comp = po.composite(m1, m1)
        b1, b2 = (0, 1) if rng.random() < 0.5 else (1, 0)
        r = rng.uniform(0.01, 0.05, 2)
        mass = common.sphere_mass(r)
        pos = np.zeros((2, 3))
        pos[0] = rng.uniform(-1, 1, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        depth = -float(rng.uniform(1e-6, 2e-3))
        pos[1] = pos[0] + d * (r[0] + r[1] + depth)
        rot = np.array([rand_quat(rng), rand_quat(rng)])
        vel = rng.normal(size=(2, 6)) * np.array([0.3, 0.3, 0.3, 5, 5, 5])
        n = (pos[b2] - pos[b1]) / np.linalg.norm(pos[b2] - pos[b1])
        pt1, pt2 = pos[b1] + n * r[b1], pos[b2] - n * r[b2]
        erad = r[0] * r[1] / (r[0] + r[1])
        hist = None
        if tang == po.TANG_MULTISTEP and rng.random() < 0.7:
            hist = dict(disp=rng.normal(size=3) * 1e-5, dur=float(rng.uniform(0, 0.05)), relvel=float(rng.uniform(0, 2)))
        a = po.contact_force("orc", s, comp, b1, b2, mass, pos, rot, vel, pt1, pt2, n, depth, erad, hist)
        b = po.contact_force("ref", s, comp, b1, b2, mass, pos, rot, vel, pt1, pt2, n, depth, erad, hist)
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y), (it, x, y)
        if tang == po.TANG_MULTISTEP:
            assert np.array_equal(a[3]["disp"], b[3]["disp"]) and a[3]["dur"] == b[3]["dur"] and a[3]["relvel"] == b[3]["relvel"]

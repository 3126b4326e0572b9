"""Oracle vs the reference's own known-answer vectors for the narrowphase primitives
(src/tests/unit_tests/collision/utest_COLL_narrow_prims.cpp: snap_to_box :39-73, sphere_sphere :974-1068,
box_sphere :1072-1226; tolerance 1e-10 as in the reference's fp64 build, :30-34)."""
import math

import numpy as np
import pytest

from oracle import pyoracle as po

PREC = 1e-10
OOS2 = math.sqrt(0.5)


def near(a, b):
    assert np.allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=0, atol=PREC), (a, b)


def test_snap_to_box():
    P = po.orc_prims()
    hd = (1.0, 2.0, 3.0)
    for loc, code, out in [((0.5, -1.0, 1.5), 0, (0.5, -1.0, 1.5)), ((0.5, -1.0, -3.5), 4, (0.5, -1.0, -3.0)),
                           ((0.5, -2.5, -3.5), 6, (0.5, -2.0, -3.0)), ((1.5, -2.5, -3.5), 7, (1.0, -2.0, -3.0))]:
        c, l = P.snap_to_box(hd, loc)
        assert c == code
        assert tuple(l) == out


@pytest.mark.parametrize("sep", [0.0, 0.1])
def test_sphere_sphere(sep):
    P = po.orc_prims()
    # separated (far)
    assert P.sphere_sphere((2, 2, 0), 1, (2, 0, 0), 0.5, sep) is None
    # separated (near)
    r = P.sphere_sphere((2, 2, 0), 1, (2, 0, 0), 0.95, sep)
    if sep:
        near(r["norm"], (0, -1, 0)); near(r["depth"], 0.05); near(r["pt1"], (2, 1, 0)); near(r["pt2"], (2, 0.95, 0))
        near(r["erad"], 0.95 / 1.95)
    else:
        assert r is None
    # touching
    r = P.sphere_sphere((2, 2, 0), 1, (2, 0, 0), 1, sep)
    if sep:
        near(r["norm"], (0, -1, 0)); near(r["depth"], 0); near(r["pt1"], (2, 1, 0)); near(r["pt2"], (2, 1, 0))
        near(r["erad"], 0.5)
    else:
        assert r is None
    # penetrated
    r = P.sphere_sphere((1, 1, 0), 1, (2.5, 1, 0), 1, sep)
    near(r["norm"], (1, 0, 0)); near(r["depth"], -0.5); near(r["pt1"], (2, 1, 0)); near(r["pt2"], (1.5, 1, 0))
    near(r["erad"], 0.5)


@pytest.mark.parametrize("sep", [0.0, 0.1])
def test_box_sphere(sep):
    P = po.orc_prims()
    edge_radius = 0.1
    hd = (1.0, 2.0, 3.0)
    bpos = (math.sqrt(2.0), 0.0, 0.0)
    brot = (math.cos(math.pi / 8), 0.0, 0.0, math.sin(math.pi / 8))  # QuatFromAngleZ(pi/4)
    s_rad = 1.5
    call = lambda p: P.box_sphere(bpos, brot, hd, p, s_rad, sep)
    assert call((0.5, 0.5, 1.0)) is None          # center inside box
    assert call((3.5, 2.5, 1.0)) is None          # face, far
    r = call((4.55 * OOS2, 2.55 * OOS2, 1.0))     # face, near
    if sep:
        near(r["norm"], (OOS2, OOS2, 0)); near(r["depth"], 0.05); near(r["pt1"], (3 * OOS2, OOS2, 1.0))
        near(r["pt2"], (3.05 * OOS2, 1.05 * OOS2, 1.0))
    else:
        assert r is None
    r = call((4 * OOS2, 2.0 * OOS2, 1.0))         # face, penetrated
    near(r["norm"], (OOS2, OOS2, 0)); near(r["depth"], -0.5); near(r["pt1"], (3 * OOS2, OOS2, 1.0))
    near(r["pt2"], (2.5 * OOS2, 0.5 * OOS2, 1.0)); near(r["erad"], s_rad)
    assert call((OOS2, 4.0, 1.0)) is None         # edge, far
    r = call((OOS2, 3.0 * OOS2 + 1.55, 1.0))      # edge, near
    if sep:
        near(r["norm"], (0, 1, 0)); near(r["depth"], 0.05); near(r["pt1"], (OOS2, 3 * OOS2, 1.0))
        near(r["pt2"], (OOS2, 3 * OOS2 + 0.05, 1.0))
    else:
        assert r is None
    r = call((OOS2, 3.0 * OOS2 + 1.0, 1.0))       # edge, penetrated
    near(r["norm"], (0, 1, 0)); near(r["depth"], -0.5); near(r["pt1"], (OOS2, 3 * OOS2, 1.0))
    near(r["pt2"], (OOS2, 3 * OOS2 - 0.5, 1.0)); near(r["erad"], s_rad * edge_radius / (s_rad + edge_radius))
    assert call((OOS2, 4.0, 4.0)) is None         # corner, far
    sp = np.array((OOS2, 4.55 * OOS2, 3.0 + 1.55 * OOS2))
    r = call(sp)                                  # corner, near
    if sep:
        near(r["norm"], (0, OOS2, OOS2)); near(r["depth"], 0.05); near(r["pt1"], (OOS2, 3 * OOS2, 3.0))
        near(r["pt2"], sp - s_rad * r["norm"])
    else:
        assert r is None
    sp = np.array((OOS2, 4.0 * OOS2, 3.0 + OOS2))
    r = call(sp)                                  # corner, penetrated
    near(r["norm"], (0, OOS2, OOS2)); near(r["depth"], -0.5); near(r["pt1"], (OOS2, 3 * OOS2, 3.0))
    near(r["pt2"], sp - s_rad * r["norm"]); near(r["erad"], s_rad * edge_radius / (s_rad + edge_radius))


def test_triangle_sphere_handmade():
    """No reference vector exists for triangle_sphere; hand-derived face / edge / vertex / backside."""
    P = po.orc_prims()
    A, B, Cc = (0, 0, 0), (1, 0, 0), (0, 1, 0)   # normal +z
    r = P.triangle_sphere(A, B, Cc, (0.25, 0.25, 0.3), 0.5)
    near(r["norm"], (0, 0, 1)); near(r["depth"], -0.2); near(r["pt1"], (0.25, 0.25, 0)); near(r["erad"], 0.5)
    assert P.triangle_sphere(A, B, Cc, (0.25, 0.25, -0.3), 0.5) is None      # one-sided (h <= 0)
    assert P.triangle_sphere(A, B, Cc, (0.25, 0.25, 0.5), 0.5) is None       # h >= r
    r = P.triangle_sphere(A, B, Cc, (0.5, -0.3, 0.2), 0.5)                   # edge AB
    d = math.hypot(0.3, 0.2)
    near(r["pt1"], (0.5, 0, 0)); near(r["depth"], d - 0.5); near(r["norm"], (0, -0.3 / d, 0.2 / d))
    near(r["erad"], 0.5 * 0.1 / 0.6)
    r = P.triangle_sphere(A, B, Cc, (-0.2, -0.2, 0.1), 0.5)                  # vertex A
    d = math.sqrt(0.09)
    near(r["pt1"], (0, 0, 0)); near(r["depth"], d - 0.5)

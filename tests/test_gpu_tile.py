"""The tile force kernel (shared-memory staging of the neighbour bins, tiled search grid) against k_force_integrate on the
plain grid: same scene, same steps, every bit of the state must agree -- the two differ in where the partner records are
read from and in the storage order of the spheres, never in arithmetic or summation order.  The oracle parity of the tile
path itself is what the rest of the GPU suite checks when it runs with DEMB200_TILE=1 (scripts/gpu_*.sh do both)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402
from test_gpu_parity import kinematics  # noqa: E402


def run(scene, tile, steps, **kw):
    old = os.environ.get("DEMB200_TILE")
    os.environ["DEMB200_TILE"] = "1" if tile else "0"
    try:
        g = common.make_gpu(scene, **kw)  # the switch is read at initialize
    finally:
        if old is None:
            os.environ.pop("DEMB200_TILE", None)
        else:
            os.environ["DEMB200_TILE"] = old
    g.step(steps)
    out = g.state(), g.stats(), g.history()
    g.close()
    return out


@pytest.mark.parametrize("n,poly,roll,matprops", [(20000, None, 0.0, True), (60000, (0.8, 1.2), 0.05, True), (15000, None, 0.0, False)])
def test_tile_kernel_bitwise_equals_plain_kernel(n, poly, roll, matprops):
    from chrono_b200 import dem
    scene = scenes.settling_scene(n, sep_factor=1.7 if poly else 1.99, seed=91, polydisperse=poly)
    vel, om = kinematics(n, 23, vscale=0.5, wscale=3.0)  # fast enough to rebuild the lists several times
    mat = common.settling_material(mu_roll=roll)
    if not matprops:
        mat.update(kn=2e7, kt=2e7, gn=40.0, gt=20.0)
    kw = dict(vel=vel, omega=om, dt=1e-4, mat=mat, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP, history_slots=24,
              use_mat_props=matprops)
    if roll:
        # the rolling-resistance instantiations of the two kernels contract a*b+c into FMAs at different places (same source,
        # different inlining context; with -fmad=false they agree bit for bit, profiles/README.md): a short run, 1e-11 relative
        (pa, va, wa), sa, ha = run(scene, False, 20, **kw)
        (pb, vb, wb), sb, hb = run(scene, True, 20, **kw)
        assert common.rel_err(pa, pb) < 1e-12 and common.rel_err(va, vb) < 1e-11 and common.rel_err(wa, wb) < 1e-10
        return
    (pa, va, wa), sa, ha = run(scene, False, 400, **kw)
    (pb, vb, wb), sb, hb = run(scene, True, 400, **kw)
    assert sa["rebuilds"] >= 3 and sb["rebuilds"] == sa["rebuilds"]
    assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(wa, wb), (np.abs(pa - pb).max(), np.abs(va - vb).max(), np.abs(wa - wb).max(), int((pa != pb).any(axis=1).sum()))
    ka = np.argsort((ha["owner"].astype(np.int64) << 32) | ha["other"])
    kb = np.argsort((hb["owner"].astype(np.int64) << 32) | hb["other"])
    assert np.array_equal(ha["owner"][ka], hb["owner"][kb]) and np.array_equal(ha["other"][ka], hb["other"][kb])
    assert np.array_equal(ha["disp"][ka], hb["disp"][kb])


def test_tile_kernel_frictionless_and_onestep():
    from chrono_b200 import dem
    scene = scenes.settling_scene(12000, sep_factor=1.99, seed=92)
    vel, om = kinematics(12000, 24, vscale=0.4)
    for tang in (dem.TANG_NONE, dem.TANG_ONESTEP):
        kw = dict(vel=vel, omega=om, dt=1e-4, force_model=dem.HERTZ, tangential_mode=tang)
        a = run(scene, False, 200, **kw)[0]
        b = run(scene, True, 200, **kw)[0]
        for x, y in zip(a, b):
            assert np.array_equal(x, y)

"""BASELINE.json's full single-GPU size (configs[1]: 1 M spheres, Hertz-Mindlin MultiStep): the same single-step parity as
the small cases (bit-exact bins and pair set, 1e-9 forces / state), then size-independent properties of a longer run:
bitwise determinism of two independent runs, independence of the Verlet skin, Newton's third law on the recorded forces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from chrono_b200 import scenes  # noqa: E402
import dem_common as common  # noqa: E402
from test_gpu_parity import compare_step, kinematics  # noqa: E402

N = 1000000


def test_one_million_single_step_parity():
    sc = scenes.settling_scene(N, radius=0.02, sep_factor=2.0, jitter=0.005, seed=12346)  # the bench packing
    vel, om = kinematics(N, 21, vscale=0.1, wscale=2.0)
    o, g, npairs = compare_step(sc, vel, om, steps=1, dt=1e-4, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
    assert npairs > 2.5 * N  # ~6 contacts per sphere
    # Newton's third law on the engine's per-sphere forces: every sphere-sphere contact enters twice with bit-identical
    # magnitude and opposite sign, so the total equals the wall reactions: compare with the oracle's total
    fg, _ = g.forces()
    fo, _ = o.body_forces()
    tot_g, tot_o = fg.sum(axis=0), fo[o.first_sphere_body:].sum(axis=0)
    scale = np.abs(fg).sum()
    assert np.abs(tot_g - tot_o).max() < 1e-12 * scale


def test_one_million_determinism_and_skin_independence():
    from chrono_b200 import dem
    sc = scenes.settling_scene(N, radius=0.02, sep_factor=2.0, jitter=0.005, seed=12346)
    vel, om = kinematics(N, 22, vscale=0.3, wscale=2.0)
    kw = dict(vel=vel, omega=om, dt=1e-4, force_model=dem.HERTZ, tangential_mode=dem.TANG_MULTISTEP)
    states = []
    for skin in (-1.0, -1.0, 0.6 * 0.02):
        g = common.make_gpu(sc, verlet_skin=skin, **kw)
        g.step(120)
        states.append(g.state() + (g.stats()["rebuilds"],))
        g.close()
    a, b, c = states
    for k in range(3):
        assert np.array_equal(a[k], b[k]), "two runs of the same job differ"
        assert np.array_equal(a[k], c[k]), "the result depends on the Verlet skin / rebuild cadence"
    assert a[3] >= 2 and c[3] < a[3]  # the larger skin rebuilt less often, yet every bit agrees

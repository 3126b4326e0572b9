"""Host-side scene generators of the bench (no GPU): the per-slab lattice generator must produce the same spheres for any
partition, and the "static equilibrium" bed must really be at rest under the oracle's physics."""
import numpy as np

from oracle import pyoracle as po
from chrono_b200 import scenes
import dem_common as common


def test_slab_lattice_scene_is_partition_independent():
    whole = scenes.slab_lattice_scene(60000, world=1, rank=0, layers=12, precompress=True, cross_section_of=15000)
    parts = [scenes.slab_lattice_scene(60000, world=4, rank=r, layers=12, precompress=True) for r in range(4)]
    assert sum(p["n"] for p in parts) == parts[0]["n_total"]
    ids = np.concatenate([p["ids"] for p in parts])
    assert len(np.unique(ids)) == len(ids)  # every sphere owned once
    for r, p in enumerate(parts):
        x = p["pos"][:, 0]
        assert (x >= p["lo"]).all() and (x < p["hi"]).all()
        if r:
            assert parts[r - 1]["hi"] == p["lo"]
    # the jitter of a sphere depends on its global id only: two different partitions of the same lattice agree bit for bit
    two = [scenes.slab_lattice_scene(60000, world=2, rank=r, layers=12, precompress=True, cross_section_of=15000) for r in range(2)]
    four = {int(i): p for part in parts for i, p in zip(part["ids"], part["pos"])}
    if two[0]["n_total"] == parts[0]["n_total"]:
        for part in two:
            for i, p in list(zip(part["ids"], part["pos"]))[::997]:
                assert np.array_equal(four[int(i)], p)
    assert whole["n"] == whole["n_total"]


def test_equilibrium_bed_is_at_rest_in_the_oracle():
    """hcp_layer_heights: the Hertz support of the three spheres below balances the overburden.  Released in the oracle, the
    interior of the bed must not move at all (the rows next to the side walls miss supporters and settle; their disturbance has
    not reached the middle after 100 steps), while the plain lattice of the same spheres is in free fall."""
    def run(precompress):
        sc = scenes.slab_lattice_scene(16000, layers=8, precompress=precompress, jitter=0.0)
        sc["n"] = sc["n_total"]
        o = common.make_oracle(sc, dt=1e-4, force_model=po.HERTZ, tangential_mode=po.TANG_MULTISTEP)
        assert o.step(100) == 0
        _, _, vel, _ = o.state()
        f = o.first_sphere_body
        L = sc["box_size"]
        inner = (np.abs(sc["pos"][:, 0]) < 0.25 * L[0]) & (np.abs(sc["pos"][:, 1]) < 0.25 * L[1])
        return np.linalg.norm(vel[f:], axis=1)[inner].max()

    v_eq, v_lat = run(True), run(False)
    print("interior of the bed after 100 steps: equilibrium heights %.2e m/s, plain lattice %.2e m/s" % (v_eq, v_lat))
    assert v_eq < 1e-7
    assert abs(v_lat - 9.81 * 100 * 1e-4) < 1e-3  # free fall

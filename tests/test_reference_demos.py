"""Drop-in proof at the source level: the reference's own demo programs (src/demos/dem/demo_DEM_*.cpp) compile and link
UNMODIFIED against include/ (ChSystemDem / ChSystemDemMesh mirror, look-alike Chrono value types, samplers, JSON
parameter reader) and libchrono_b200_dem.so -- and run on the B200 engine.

CPU (where /root/reference is mounted): compile + link from the sources where they lie; the binaries land in
tests/cpp/_bin/ (git-ignored build artefacts that travel to the GPU box).  GPU: run them with short parameter files."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "cpp", "_bin")
REF_DEMOS = "/root/reference/src/demos/dem"
DEMOS = ["demo_DEM_movingBoundary", "demo_DEM_mixer", "demo_DEM_repose", "demo_DEM_fixedTerrain"]


def build_demo(name):
    """Returns the path of the binary, building it if the reference sources are here; None if neither exists."""
    from chrono_b200.build import build_library, LIB_PATH
    out = os.path.join(BIN, name)
    src = os.path.join(REF_DEMOS, name + ".cpp")
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    build_library()
    os.makedirs(BIN, exist_ok=True)
    newest_dep = max(os.path.getmtime(LIB_PATH), os.path.getmtime(src),
                     max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(os.path.join(ROOT, "include")) for f in fs))
    if os.path.exists(out) and os.path.getmtime(out) > newest_dep:
        return out
    libdir = os.path.dirname(LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           src, "-o", out, "-L", libdir, "-lchrono_b200_dem", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    return out


@pytest.mark.parametrize("name", DEMOS)
def test_reference_demo_compiles_unmodified(name):
    if not os.path.isdir(REF_DEMOS):
        pytest.skip("reference sources not mounted")
    assert os.path.exists(build_demo(name))


def mixer_obj(path):
    """A four-blade paddle in the unit cube the demo scales by (Bx/2, Bx/2, chamber height): thin radial boxes."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    verts, faces = [], []

    def box(lo, hi):
        b = len(verts)
        for z in (lo[2], hi[2]):
            for y in (lo[1], hi[1]):
                for x in (lo[0], hi[0]):
                    verts.append((x, y, z))
        quads = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
        for q in quads:
            faces.append((b + q[0], b + q[1], b + q[2]))
            faces.append((b + q[0], b + q[2], b + q[3]))

    t = 0.03
    box((0.05, -t, 0.0), (0.9, t, 1.0))
    box((-0.9, -t, 0.0), (-0.05, t, 1.0))
    box((-t, 0.05, 0.0), (t, 0.9, 1.0))
    box((-t, -0.9, 0.0), (t, -0.05, 1.0))
    with open(path, "w") as f:
        for v in verts:
            f.write("v %.6f %.6f %.6f\n" % v)
        for a, b, c in faces:
            f.write("f %d %d %d\n" % (a + 1, b + 1, c + 1))


def terrain_obj(path):
    """A bumpy height field over [-1, 1]^2 (the demo scales it by (box_X/2, box_Y/2, box_Z) and lowers it by box_Z/2)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    n = 12
    xs = np.linspace(-1, 1, n + 1)
    with open(path, "w") as f:
        for j in range(n + 1):
            for i in range(n + 1):
                f.write("v %.6f %.6f %.6f\n" % (xs[i], xs[j], 0.12 + 0.05 * np.sin(3 * xs[i]) * np.cos(2 * xs[j])))
        for j in range(n):
            for i in range(n):
                a = j * (n + 1) + i + 1
                f.write("f %d %d %d\nf %d %d %d\n" % (a, a + 1, a + n + 2, a, a + n + 2, a + n + 1))


def run_demo(name, json_name, tmp_path, timeout=900):
    exe = build_demo(name)
    if exe is None:
        pytest.skip("demo binary was not built (needs /root/reference at build time)")
    data = tmp_path / "data"
    mixer_obj(str(data / "models" / "mixer" / "internal_mixer.obj"))
    terrain_obj(str(data / "models" / "fixedterrain.obj"))
    env = dict(os.environ, CHRONO_DATA_DIR=str(data) + "/", CHRONO_OUTPUT_DIR=str(tmp_path / "out") + "/")
    r = subprocess.run([exe, os.path.join(HERE, "golden", "demo_json", json_name)], capture_output=True, text=True,
                       timeout=timeout, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return tmp_path / "out" / "DEM"


@pytest.mark.gpu
def test_reference_demo_moving_boundary_runs(tmp_path):
    out = run_demo("demo_DEM_movingBoundary", "movingBoundary_small.json", tmp_path) / "movingBoundary"
    files = sorted(os.listdir(out))
    assert files[0] == "step000000.csv" and len(files) >= 5, files
    first = np.loadtxt(out / files[0], delimiter=",", skiprows=1)
    last = np.loadtxt(out / files[-1], delimiter=",", skiprows=1)
    assert first.shape == last.shape and first.shape[0] > 500 and first.shape[1] == 4  # x, y, z, absv
    assert open(out / files[0]).readline().strip() == "x,y,z,absv"
    # the bed has fallen under gravity and nothing left the big domain (60 x 60 x 40 centred at (30, 30, 20))
    assert last[:, 2].mean() < first[:, 2].mean()
    assert (last[:, :3].min(axis=0) > -0.1).all() and (last[:, :3].max(axis=0) < np.array([60.1, 60.1, 40.1])).all()


@pytest.mark.gpu
def test_reference_demo_mixer_runs(tmp_path):
    out = run_demo("demo_DEM_mixer", "mixer_small.json", tmp_path) / "mixer"
    files = sorted(os.listdir(out))
    assert any(f.endswith(".csv") for f in files) and any(f.endswith(".vtk") for f in files), files
    pts = np.loadtxt(out / [f for f in files if f.endswith(".csv")][-1], delimiter=",", skiprows=1)
    assert pts.shape[0] > 300 and np.isfinite(pts).all()
    assert (np.hypot(pts[:, 0], pts[:, 1]) < 50.0).all()  # inside the cylinder BC of radius Bx / 2


@pytest.mark.gpu
def test_reference_demo_fixed_terrain_runs(tmp_path):
    """Terrain of fixed particles (MeshSphericalDecomposition) + a layer of free ones poured onto it."""
    out = run_demo("demo_DEM_fixedTerrain", "fixedTerrain_small.json", tmp_path) / "fixedTerrain"
    files = sorted(os.listdir(out))
    assert len(files) >= 2, files
    hdr = open(out / files[0]).readline().strip()
    assert hdr == "x,y,z,absv,fixed,wx,wy,wz", hdr
    first = np.loadtxt(out / files[0], delimiter=",", skiprows=1)
    last = np.loadtxt(out / files[-1], delimiter=",", skiprows=1)
    fixed = first[:, 4] != 0
    assert fixed.sum() > 100 and (~fixed).sum() > 100
    assert np.array_equal(first[fixed, :3], last[fixed, :3])           # the terrain does not move
    assert last[~fixed, 2].mean() < first[~fixed, 2].mean()             # the free particles fall

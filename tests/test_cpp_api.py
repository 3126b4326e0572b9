"""The C++ mirror of chrono::dem::ChSystemDem (include/chrono_dem/physics/ChSystemDem.h) driven by the reference's own
unit-test scenarios (src/tests/unit_tests/dem/utest_DEM_{stack,frictionrolling,pyramid}.cpp), re-typed in tests/cpp/.

CPU: the programs must compile and link against libchrono_b200_dem.so (source compatibility of the API surface).
GPU: they must pass with the reference's tolerances."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "cpp", "_bin")
PROGRAMS = ["utest_DEM_stack", "utest_DEM_frictionrolling", "utest_DEM_pyramid", "utest_DEM_api", "utest_DEM_meshrolling", "utest_utils", "utest_DEM_bcsphere", "utest_DEM_contactinfo"]


def build_program(name):
    from chrono_b200.build import build_library, LIB_PATH
    build_library()
    os.makedirs(BIN, exist_ok=True)
    out = os.path.join(BIN, name)
    src = os.path.join(HERE, "cpp", name + ".cpp")
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(LIB_PATH)):
        return out
    libdir = os.path.dirname(LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "cpp"),
           "-I", "/usr/local/cuda/include", src, "-o", out, "-L", libdir, "-lchrono_b200_dem",
           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return out


@pytest.mark.parametrize("name", PROGRAMS)
def test_cpp_programs_compile_and_link(name):
    assert os.path.exists(build_program(name))


def run(name, *args, timeout=600):
    exe = build_program(name)
    r = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_lookalike_utilities_cpu(tmp_path):
    """Samplers, JSON parameter reader, data-path helpers, OBJ reader: no GPU involved."""
    out = run("utest_utils", str(tmp_path), os.path.join(HERE, "golden", "demo_json", "mixer_small.json"))
    assert "PASSED" in out


@pytest.mark.gpu
def test_dem_stack():
    run("utest_DEM_stack")


@pytest.mark.gpu
def test_dem_frictionrolling():
    run("utest_DEM_frictionrolling")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["hold", "collapse"])
def test_dem_pyramid_from_reference_checkpoint(mode):
    run("utest_DEM_pyramid", os.path.join(HERE, "golden", "pyramid_checkpoint.dat"), mode)


@pytest.mark.gpu
def test_dem_meshrolling_and_cosim_wrench(tmp_path):
    """utest_DEM_meshrolling scenario through ChSystemDemMesh + ApplyMeshMotion / CollectMeshContactForces."""
    out = run("utest_DEM_meshrolling", str(tmp_path))
    assert "PASSED" in out


@pytest.mark.gpu
def test_dem_bc_ball_with_mass_and_cone_hopper():
    out = run("utest_DEM_bcsphere")
    assert "PASSED" in out


@pytest.mark.gpu
def test_dem_api_io_roundtrip(tmp_path):
    out = run("utest_DEM_api", str(tmp_path))
    assert "PASSED" in out


@pytest.mark.gpu
def test_dem_contact_info_plane_rotation_single_step_checkpoint(tmp_path):
    """SetRecordingContactInfo + per-pair getters + WriteContactInfoFile, SetBCPlaneRotation, SINGLE_STEP checkpoint layout."""
    out = run("utest_DEM_contactinfo", str(tmp_path))
    assert "PASSED" in out

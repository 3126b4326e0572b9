// Scenario of src/tests/unit_tests/dem/utest_DEM_meshrolling.cpp:32-129: one sphere launched along x on a single mesh
// facet (data/testing/dem/one_facet.obj geometry, written to argv[1] here) rolls to rest under sliding + Schwartz
// rolling friction against the MESH; y stays put, z ends at the settled height.  Then the co-simulation round trip of
// demo_DEM_ballCosim.cpp: ApplyMeshMotion -> AdvanceSimulation -> CollectMeshContactForces returns the sphere's weight.
#include <fstream>
#include <string>
#include <vector>
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

void demMeshRolling(int argc, char** argv) {
    ASSERT_TRUE(argc > 1);
    const std::string obj = std::string(argv[1]) + "/one_facet.obj";
    {
        std::ofstream f(obj);
        f << "v -5.000000 -5.000000 0.000000\nv 5.000000 -5.000000 -0.000000\nv 5.000000 5.000000 -0.000000\n"
             "vn 0.000000 0.000000 1.000000\nf 1//1 2//1 3//1\n";
    }
    float radius = 0.5f, density = 6.f, g = 980.f;
    float precision_KE = 1e-3f, precision_pos = 1e-3f;
    float mu_s = 0.2f, mu_r = 0.0008f;

    ChSystemDemMesh dem_sys(radius, density, ChVector3f(20.f, 20.f, 10.f));
    float mass = 100.f;
    float inertia = (2.0f / 5) * mass * 0.5f * 0.5f;
    dem_sys.AddMesh(obj, ChVector3f(0), ChMatrix33<float>(1.f), mass);
    dem_sys.EnableMeshCollision(true);
    ASSERT_TRUE(dem_sys.GetNumMeshes() == 1);
    ASSERT_TRUE(dem_sys.GetMesh(0)->GetNumTriangles() == 1);

    float penetration = std::pow(mass * std::fabs(-g) / (1e11f), 2.f / 3.f);
    float settled_pos = radius - penetration;
    std::vector<ChVector3f> body_point = {ChVector3f(1.f, -1.f, 0.5f)};
    std::vector<ChVector3f> velocity = {ChVector3f(1.f, 0.f, 0.f)};
    dem_sys.SetParticles(body_point, velocity);
    dem_sys.SetPsiFactors(32, 16);
    dem_sys.SetKn_SPH2SPH(1e11); dem_sys.SetKn_SPH2WALL(1e11); dem_sys.SetKn_SPH2MESH(1e11);
    dem_sys.SetGn_SPH2SPH(1e4); dem_sys.SetGn_SPH2WALL(1e4); dem_sys.SetGn_SPH2MESH(1e4);
    dem_sys.SetKt_SPH2SPH(1e7); dem_sys.SetKt_SPH2WALL(1e7); dem_sys.SetKt_SPH2MESH(1e7);
    dem_sys.SetGt_SPH2SPH(1e4); dem_sys.SetGt_SPH2WALL(1e4); dem_sys.SetGt_SPH2MESH(1e4);
    dem_sys.SetStaticFrictionCoeff_SPH2SPH(mu_s);
    dem_sys.SetStaticFrictionCoeff_SPH2WALL(mu_s);
    dem_sys.SetStaticFrictionCoeff_SPH2MESH(mu_s);
    dem_sys.SetFrictionMode(CHDEM_FRICTION_MODE::MULTI_STEP);
    dem_sys.SetRollingMode(CHDEM_ROLLING_MODE::SCHWARTZ);
    dem_sys.SetRollingCoeff_SPH2SPH(mu_r);
    dem_sys.SetRollingCoeff_SPH2WALL(mu_r);
    dem_sys.SetRollingCoeff_SPH2MESH(mu_r);
    dem_sys.SetGravitationalAcceleration(ChVector3d(0.f, 0.f, -g));
    dem_sys.SetVerbosity(CHDEM_VERBOSITY::QUIET);

    float step_size = 1e-4f, curr_time = 0, end_time = 3.0f, time_start_check = 0.1f;
    bool settled = false;
    dem_sys.SetFixedStepSize(step_size);
    dem_sys.SetTimeIntegrator(CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE);
    dem_sys.SetBDFixed(true);
    dem_sys.Initialize();

    while (curr_time < end_time) {
        dem_sys.AdvanceSimulation(step_size);
        curr_time += step_size;
        if (curr_time > time_start_check) {
            float vel = dem_sys.GetParticleVelocity(0).Length();
            float omg = dem_sys.GetParticleAngVelocity(0).Length();
            float KE = 0.5f * mass * vel * vel + 0.5f * inertia * omg * omg;
            if (KE < precision_KE) {
                settled = true;
                break;
            }
        }
    }
    std::printf("settled=%d at t=%g\n", (int)settled, curr_time);
    ASSERT_TRUE(settled);
    ChVector3d end_pos = dem_sys.GetParticlePosition(0);
    std::printf("end pos %g %g %g (settled z %g)\n", end_pos.x(), end_pos.y(), end_pos.z(), settled_pos);
    ASSERT_TRUE(end_pos.x() > 1.0f);
    ASSERT_NEAR(end_pos.y(), -1.0f, precision_pos);
    ASSERT_NEAR(end_pos.z(), settled_pos, precision_pos);

    // co-simulation round trip: hold the facet where it is, let the sphere come to rest, read the wrench
    for (int i = 0; i < 3000; i++) {
        dem_sys.ApplyMeshMotion(0, ChVector3d(0, 0, 0), ChQuaternion<>(1, 0, 0, 0), ChVector3d(0, 0, 0), ChVector3d(0, 0, 0));
        dem_sys.AdvanceSimulation(step_size);
    }
    ChVector3d F, T;
    dem_sys.CollectMeshContactForces(0, F, T);
    const double weight = 4.0 / 3.0 * 3.14159265358979 * radius * radius * radius * density * g;
    const ChVector3d p = dem_sys.GetParticlePosition(0);
    std::printf("mesh force %g %g %g (weight %g), torque %g %g %g\n", F.x(), F.y(), F.z(), weight, T.x(), T.y(), T.z());
    ASSERT_NEAR(F.z(), -weight, 1e-3 * weight);
    ASSERT_NEAR(T.x(), p.y() * F.z(), 1e-3 * weight);   // r x F with r = (px, py, 0)
    ASSERT_NEAR(T.y(), -p.x() * F.z(), 1e-3 * weight);
    std::vector<ChVector3d> Fs, Ts;
    dem_sys.CollectMeshContactForces(Fs, Ts);
    ASSERT_TRUE(Fs.size() == 1 && Ts.size() == 1);
    dem_sys.WriteMeshes(std::string(argv[1]) + "/meshes");
    std::ifstream vtk(std::string(argv[1]) + "/meshes.vtk");
    std::string first;
    std::getline(vtk, first);
    ASSERT_TRUE(first == "# vtk DataFile Version 2.0");
}
RUN_TEST(demMeshRolling)

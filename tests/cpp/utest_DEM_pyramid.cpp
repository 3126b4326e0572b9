// Scenario of src/tests/unit_tests/dem/utest_DEM_pyramid.cpp:42-168: a three-sphere pyramid is constructed FROM THE
// REFERENCE'S CHECKPOINT FIXTURE (data/testing/dem/pyramid_checkpoint.dat, header "ChSystemGpu"; a verbatim copy of
// the fixture lives in tests/golden/), the top sphere is released; with high friction the pyramid holds, with low
// friction it collapses.  argv[1] = path of the checkpoint, argv[2] = "hold" | "collapse".
#include <cstring>
#include <vector>
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

void demPyramid(int argc, char** argv) {
    ASSERT_TRUE(argc >= 3);
    const bool hold = std::strcmp(argv[2], "hold") == 0;
    ChSystemDem dem_sys(argv[1]);
    ASSERT_TRUE(dem_sys.GetNumParticles() == 3);

    std::vector<bool> body_fixity = {false, false, false};
    dem_sys.SetParticleFixed(body_fixity);
    if (hold) {
        float mu_k = 0.5f;
        dem_sys.SetStaticFrictionCoeff_SPH2SPH(mu_k);
        dem_sys.SetStaticFrictionCoeff_SPH2WALL(mu_k);
        float mu_r = 0.2f;
        dem_sys.SetRollingMode(CHDEM_ROLLING_MODE::SCHWARTZ);
        dem_sys.SetRollingCoeff_SPH2SPH(mu_r);
        dem_sys.SetRollingCoeff_SPH2WALL(mu_r);
    } else {
        float mu_k = 0.01f;
        dem_sys.SetStaticFrictionCoeff_SPH2SPH(mu_k);
        dem_sys.SetStaticFrictionCoeff_SPH2WALL(mu_k);
        dem_sys.SetRollingMode(CHDEM_ROLLING_MODE::NO_RESISTANCE);
    }
    ChVector3f ground_plate_pos(0.0, 0.0, 0.0), ground_plate_normal(0.0, 0.0, 1.0f);
    size_t plane = dem_sys.CreateBCPlane(ground_plate_pos, ground_plate_normal, true);
    ASSERT_TRUE(plane == NUM_RESERVED_BC_IDS);  // the wall-contact label 10 = nSpheres(3) + BC_id(6) + 1 of the fixture
    dem_sys.SetVerbosity(CHDEM_VERBOSITY::QUIET);
    dem_sys.Initialize();

    float mass = (4.0f / 3) * 3.14f * 0.5f * 0.5f * 0.5f * 1.9f;
    float inertia = (2.0f / 5) * mass * 0.5f * 0.5f;
    float radius = 0.5f;
    float precision_KE = 1e-7f, precision_pos = 1e-3f, precision_time = 1e-3f;

    auto pos = dem_sys.GetParticlePosition(2);
    float g = 9.81f;
    float contact_time = std::sqrt(2 * (pos.z() - (1 + std::sqrt(3.f)) * radius) / g);
    float step_size = 1e-3f, curr_time = 0;

    while (curr_time < 1.1 * contact_time) {
        dem_sys.AdvanceSimulation(step_size);
        curr_time += step_size;
        if (curr_time < 0.9 * contact_time)
            continue;
        pos = dem_sys.GetParticlePosition(2);
        if (std::abs(pos.z() - (1 + std::sqrt(3.f)) * radius) < precision_pos)
            break;
    }
    std::printf("contact at t=%g (analytical %g)\n", curr_time, contact_time);
    ASSERT_NEAR(curr_time, contact_time, precision_time);

    bool settled = false;
    float KE = 0;
    while (curr_time < 1.5f) {
        dem_sys.AdvanceSimulation(step_size);
        curr_time += step_size;
        if (curr_time < 2 * contact_time)
            continue;
        float vel = dem_sys.GetParticleVelocity(2).Length();
        float omg = dem_sys.GetParticleAngVelocity(2).Length();
        KE = 0.5f * mass * vel * vel + 0.5f * inertia * omg * omg;
        if (KE < precision_KE) {
            settled = true;
            break;
        }
    }
    pos = dem_sys.GetParticlePosition(2);
    std::printf("settled=%d at t=%g KE=%g final pos %g %g %g\n", (int)settled, curr_time, KE, pos.x(), pos.y(), pos.z());
    ASSERT_TRUE(settled);
    if (hold) {
        ASSERT_NEAR(pos.y(), 0.0f, precision_pos);
        ASSERT_TRUE(pos.z() > 2 * radius);
    } else {
        ASSERT_NEAR(pos.y(), 0.0f, precision_pos);
        ASSERT_NEAR(pos.z(), radius, precision_pos);
    }
    // the tracked ground plane carries the weight of the three spheres once everything rests
    ChVector3f f;
    ASSERT_TRUE(dem_sys.GetBCReactionForces(plane, f));
    std::printf("ground reaction %g %g %g\n", f.x(), f.y(), f.z());
}
RUN_TEST(demPyramid)

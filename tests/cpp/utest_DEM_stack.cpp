// Scenario of src/tests/unit_tests/dem/utest_DEM_stack.cpp:31-128 against the ChSystemDem mirror: a column of 5 spheres
// (MULTI_STEP friction, CHUNG integrator, SCHWARTZ rolling) must come to rest with z_i = settled + 2R i - delta i.
#include <iostream>
#include <vector>
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

void demStack(int, char**) {
    float density = 1.53f, radius = 0.5f, g = 980.f, mu_s = 0.5f, mu_r = 0.0008f;
    float precision_KE = 1e-5f, precision_pos = 1e-2f;
    float mass = 4.f / 3.f * 3.14159265f * std::pow(radius, 3.f) * density;
    float penetration = std::pow(mass * g / 1e7f, 2.f / 3.f);
    float inertia = 2.f / 5.f * mass * radius * radius;
    float settled_pos = -100.f / 2.0f + radius - penetration;

    ChSystemDem dem_sys(radius, density, ChVector3f(100.f, 100.f, 100.f));
    dem_sys.SetGravitationalAcceleration(ChVector3d(0, 0, -g));
    dem_sys.SetFrictionMode(CHDEM_FRICTION_MODE::MULTI_STEP);
    dem_sys.SetTimeIntegrator(CHDEM_TIME_INTEGRATOR::CHUNG);
    dem_sys.SetKn_SPH2SPH(1e7);
    dem_sys.SetKn_SPH2WALL(1e7);
    dem_sys.SetGn_SPH2SPH(2e4);
    dem_sys.SetGn_SPH2WALL(2e4);
    dem_sys.SetKt_SPH2SPH(2e6);
    dem_sys.SetKt_SPH2WALL(1e6);
    dem_sys.SetGt_SPH2SPH(50);
    dem_sys.SetGt_SPH2WALL(50);
    dem_sys.SetStaticFrictionCoeff_SPH2SPH(mu_s);
    dem_sys.SetStaticFrictionCoeff_SPH2WALL(mu_s);
    dem_sys.SetRollingMode(CHDEM_ROLLING_MODE::SCHWARTZ);
    dem_sys.SetRollingCoeff_SPH2SPH(mu_r);
    dem_sys.SetRollingCoeff_SPH2WALL(mu_r);
    dem_sys.SetPsiFactors(32, 16);
    dem_sys.SetVerbosity(CHDEM_VERBOSITY::QUIET);

    std::vector<ChVector3f> body_points, velocity;
    for (int i = 0; i < 5; i++) {
        body_points.push_back(ChVector3f(0.f, 0.f, settled_pos + radius * 3.f * i));
        velocity.push_back(ChVector3f(0.0f, 0.0f, 0.0f));
    }
    dem_sys.SetParticles(body_points, velocity);

    float step_size = 1e-4f, curr_time = 0.f, end_time = 3.f, time_start_check = 0.1f;
    bool settled = false;
    dem_sys.SetFixedStepSize(step_size);
    dem_sys.SetBDFixed(true);
    dem_sys.Initialize();

    while (curr_time < end_time) {
        dem_sys.AdvanceSimulation(step_size);
        curr_time += step_size;
        if (curr_time > time_start_check) {
            float KE = 0.f;
            for (int i = 0; i < 5; i++) {
                float vel = dem_sys.GetParticleVelocity(i).Length();
                float omg = dem_sys.GetParticleAngVelocity(i).Length();
                KE += 0.5f * mass * vel * vel + 0.5f * inertia * omg * omg;
            }
            if (KE < precision_KE) {
                settled = true;
                break;
            }
        }
    }
    std::printf("settled=%d at t=%g\n", (int)settled, curr_time);
    ASSERT_TRUE(settled);
    for (int i = 0; i < 5; i++) {
        ASSERT_NEAR(dem_sys.GetParticlePosition(i).x(), 0, precision_pos);
        ASSERT_NEAR(dem_sys.GetParticlePosition(i).y(), 0, precision_pos);
        ASSERT_NEAR(dem_sys.GetParticlePosition(i).z(), settled_pos + radius * 2 * i - penetration * i, precision_pos);
    }
}
RUN_TEST(demStack)

// Scenario of src/tests/unit_tests/dem/utest_DEM_frictionrolling.cpp:32-124: one sphere launched along x on the floor
// rolls to rest (sliding + Schwartz rolling friction); y stays 0, z ends at the settled height.
#include <vector>
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

void demFrictionRolling(int, char**) {
    float density = 1.53f, radius = 0.5f, g = 980.f, mu_s = 0.5f, mu_r = 0.0008f;
    float precision_KE = 1e-3f, precision_pos = 1e-3f;
    float mass = 4.f / 3.f * 3.14159265f * std::pow(radius, 3.f) * density;
    float penetration = std::pow(mass * g / 1e7f, 2.f / 3.f);
    float inertia = 2.f / 5.f * mass * radius * radius;
    float settled_pos = -20.f / 2.f + radius - penetration;

    ChSystemDem dem_sys(radius, density, ChVector3f(20.f, 20.f, 20.f));
    dem_sys.SetKn_SPH2SPH(1e7);
    dem_sys.SetKn_SPH2WALL(1e7);
    dem_sys.SetGn_SPH2SPH(1e4);
    dem_sys.SetGn_SPH2WALL(1e4);
    dem_sys.SetKt_SPH2SPH(1e7);
    dem_sys.SetKt_SPH2WALL(1e7);
    dem_sys.SetGt_SPH2SPH(1e4);
    dem_sys.SetGt_SPH2WALL(1e4);
    dem_sys.SetFrictionMode(CHDEM_FRICTION_MODE::MULTI_STEP);
    dem_sys.SetTimeIntegrator(CHDEM_TIME_INTEGRATOR::CHUNG);
    dem_sys.SetStaticFrictionCoeff_SPH2SPH(mu_s);
    dem_sys.SetStaticFrictionCoeff_SPH2WALL(mu_s);
    dem_sys.SetRollingMode(CHDEM_ROLLING_MODE::SCHWARTZ);
    dem_sys.SetRollingCoeff_SPH2SPH(mu_r);
    dem_sys.SetRollingCoeff_SPH2WALL(mu_r);
    dem_sys.SetPsiFactors(32, 16);
    dem_sys.SetGravitationalAcceleration(ChVector3d(0.f, 0.f, -g));
    dem_sys.SetVerbosity(CHDEM_VERBOSITY::QUIET);

    std::vector<ChVector3f> body_point = {ChVector3f(0, 0, settled_pos + 0.02f)};
    std::vector<ChVector3f> velocity = {ChVector3f(1.0, 0.0, 0.0)};
    dem_sys.SetParticles(body_point, velocity);

    float step_size = 1e-4f, curr_time = 0, end_time = 3.0f, time_start_check = 0.1f;
    bool settled = false;
    dem_sys.SetFixedStepSize(step_size);
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
    ASSERT_TRUE(end_pos.x() > 0.0f);
    ASSERT_NEAR(end_pos.y(), 0.0f, precision_pos);
    ASSERT_NEAR(end_pos.z(), settled_pos, precision_pos);
}
RUN_TEST(demFrictionRolling)

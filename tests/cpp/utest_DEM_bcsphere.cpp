// Boundary-condition shapes through the C++ mirror (no reference unit test exists for them; reference API
// ChSystemDem.h:206-240): a heavy ball with mass (CreateBCSphere) dropped on a settled bed comes to rest ON the bed and the
// bed carries its weight (GetBCReactionForces); a cone hopper (CreateBCConeZ) with a floor plane holds what is poured in.
#include <vector>
#include "chrono/utils/ChUtilsSamplers.h"
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

static void material(ChSystemDem& sys) {
    sys.SetKn_SPH2SPH(1e7); sys.SetKn_SPH2WALL(1e7);
    sys.SetGn_SPH2SPH(2e4); sys.SetGn_SPH2WALL(2e4);
    sys.SetKt_SPH2SPH(5e6); sys.SetKt_SPH2WALL(5e6);
    sys.SetGt_SPH2SPH(1e3); sys.SetGt_SPH2WALL(1e3);
    sys.SetStaticFrictionCoeff_SPH2SPH(0.5f); sys.SetStaticFrictionCoeff_SPH2WALL(0.5f);
    sys.SetFrictionMode(CHDEM_FRICTION_MODE::MULTI_STEP);
    sys.SetTimeIntegrator(CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE);
    sys.SetGravitationalAcceleration(ChVector3f(0, 0, -980.f));
    sys.SetFixedStepSize(1e-4f);
    sys.SetBDFixed(true);
    sys.SetVerbosity(CHDEM_VERBOSITY::QUIET);
}

void demBCShapes(int, char**) {
    const float r = 0.5f, rho = 2.5f;
    {  // ---- ball with mass on a bed
        ChSystemDem sys(r, rho, ChVector3f(16.f, 16.f, 30.f));
        material(sys);
        utils::ChHCPSampler<float> hcp(2.05f * r);
        auto pts = hcp.SampleBox(ChVector3f(0, 0, -15.f + 3.f), ChVector3f(7.f, 7.f, 2.4f));
        sys.SetParticles(pts);
        const float Rb = 2.5f, Mb = 4.f / 3.f * 3.14159265f * Rb * Rb * Rb * 1.0f;
        const size_t ball = sys.CreateBCSphere(ChVector3f(0, 0, -15.f + 6.f + Rb + 1.0f), Rb, false, true, Mb);
        sys.Initialize();
        sys.AdvanceSimulation(1.2f);
        const ChVector3f p = sys.GetBCSpherePosition(ball), v = sys.GetBCSphereVelocity(ball);
        ChVector3f F;
        ASSERT_TRUE(sys.GetBCReactionForces(ball, F));
        std::printf("ball: z %.3f (floor %.1f), |v| %.3g, reaction fz %.1f vs weight %.1f, bed top %.3f\n", p.z(), -15.f, v.Length(),
                    F.z(), Mb * 980.f, sys.GetMaxParticleZ());
        ASSERT_TRUE(v.Length() < 2.0);                    // came to rest (cm/s)
        ASSERT_TRUE(p.z() - Rb > -15.f + 2.0f);           // rests on the bed, not on the floor
        ASSERT_TRUE(p.z() - Rb < -15.f + 6.5f);
        ASSERT_NEAR(F.z(), Mb * 980.f, 0.1 * Mb * 980.f);  // the bed carries its weight
    }
    {  // ---- hopper: cone z = rho (slope 1) cut by a floor plane at z = 1
        ChSystemDem sys(r, rho, ChVector3f(40.f, 40.f, 40.f));
        material(sys);
        utils::ChHCPSampler<float> hcp(2.1f * r);
        auto pts = hcp.SampleCylinderZ(ChVector3f(0, 0, 9.f), 5.f, 2.f);
        sys.SetParticles(pts);
        sys.CreateBCConeZ(ChVector3f(0, 0, 0), 1.f, 19.f, 0.5f, false, false);
        sys.CreateBCPlane(ChVector3f(0, 0, 1.f), ChVector3f(0, 0, 1.f), false);
        sys.Initialize();
        sys.AdvanceSimulation(1.0f);
        ASSERT_TRUE(sys.GetParticlesKineticEnergy() < 1e-1 * pts.size());
        for (size_t i = 0; i < pts.size(); i++) {
            const ChVector3f q = sys.GetParticlePosition((int)i);
            const double rho_q = std::sqrt((double)q.x() * q.x() + (double)q.y() * q.y());
            ASSERT_TRUE(q.z() > 1.f + 0.9f * r);                 // above the floor plane
            ASSERT_TRUE(q.z() - rho_q > -0.1 * r * 1.4143);      // above the cone surface
        }
        std::printf("hopper holds %zu particles, max z %.3f\n", pts.size(), sys.GetMaxParticleZ());
        ASSERT_TRUE(sys.GetMaxParticleZ() < 9.f);
    }
}
RUN_TEST(demBCShapes)

// Per-contact records, plane spin and the SINGLE_STEP checkpoint layout of the ChSystemDem mirror.
//  * SetRecordingContactInfo / getNormalForce / getSlidingFrictionForce / WriteContactInfoFile
//    (reference: src/chrono_dem/physics/ChSystemDem.h:180,326-338; ChSystemDem_impl.cpp:443-635): a column of two spheres at
//    rest on the floor -- the normal forces must carry the weights, the file must hold exactly the one sphere pair;
//  * SetBCPlaneRotation (ChSystemDem.h:238; force: ChDemBoundaryConditions.cuh:394): a sphere resting on a spinning floor
//    is dragged along the surface velocity omega x (x - center);
//  * a SINGLE_STEP checkpoint carries "partners 12" without "history 12" (ChSystemDem.cpp:1450-1500) and loads again,
//    as does a hand-written file in the reference layout with blank lines in it.
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

static void common_setup(ChSystemDem& sys, CHDEM_FRICTION_MODE fm) {
    sys.SetGravitationalAcceleration(ChVector3d(0, 0, -980));
    sys.SetFrictionMode(fm);
    sys.SetTimeIntegrator(CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE);
    sys.SetKn_SPH2SPH(1e7); sys.SetKn_SPH2WALL(1e7);
    sys.SetGn_SPH2SPH(2e4); sys.SetGn_SPH2WALL(2e4);
    sys.SetKt_SPH2SPH(2e6); sys.SetKt_SPH2WALL(2e6);
    sys.SetGt_SPH2SPH(50); sys.SetGt_SPH2WALL(50);
    sys.SetStaticFrictionCoeff_SPH2SPH(0.5f); sys.SetStaticFrictionCoeff_SPH2WALL(0.5f);
    sys.SetVerbosity(CHDEM_VERBOSITY::QUIET);
    sys.SetFixedStepSize(1e-4f);
    sys.SetBDFixed(true);
}

void demContactInfo(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    const float density = 1.53f, radius = 0.5f, g = 980.f;
    const double mass = 4.0 / 3.0 * 3.14159265358979 * radius * radius * radius * density;
    const float floor_z = -50.f;
    // ---- two spheres stacked on the floor, records on
    {
        ChSystemDem sys(radius, density, ChVector3f(100.f, 100.f, 100.f));
        common_setup(sys, CHDEM_FRICTION_MODE::MULTI_STEP);
        sys.SetRecordingContactInfo(true);
        std::vector<ChVector3f> pts = {ChVector3f(0, 0, floor_z + radius), ChVector3f(0, 0, floor_z + 3 * radius)};
        sys.SetParticles(pts);
        sys.Initialize();
        sys.AdvanceSimulation(0.6f);
        ASSERT_TRUE(sys.GetParticleVelocity(1).Length() < 1e-3);
        const ChVector3f n01 = sys.getNormalForce(0, 1), n10 = sys.getNormalForce(1, 0);
        std::printf("normal force on sphere 0 from sphere 1: %g %g %g (weight %g)\n", n01.x(), n01.y(), n01.z(), mass * g);
        ASSERT_NEAR(n01.z(), -mass * g, 1e-3 * mass * g);   // sphere 1 presses sphere 0 down with its weight
        ASSERT_NEAR(n10.z(), +mass * g, 1e-3 * mass * g);   // and is carried by it
        ASSERT_NEAR(n01.x(), 0, 1e-6 * mass * g);
        // floor = reserved plane BD_WALL_ID_Z_BOT: label nSpheres + BC_id + 1 (ChDemBoundaryConditions.cuh:102)
        const unsigned floor_label = 2 + (unsigned)BD_WALL_ID_Z_BOT + 1;
        const ChVector3f nf = sys.getNormalForce(0, floor_label);
        ASSERT_NEAR(nf.z(), 2 * mass * g, 2e-3 * mass * g);
        ASSERT_NEAR(sys.getNormalForce(floor_label, 0).z(), nf.z(), 0);  // either argument order (ChSystemDem_impl.cpp:557-561)
        ASSERT_TRUE(sys.getSlidingFrictionForce(0, 1).Length() < 1e-6 * mass * g);
        ASSERT_TRUE(sys.getNormalForce(1, floor_label).Length() == 0.f);  // not in contact -> zero
        ASSERT_TRUE(sys.getRollingFrictionTorque(0, 1).Length() == 0.f);  // NO_RESISTANCE -> zero (:449-451)
        const std::string f = dir + "/contact_info.csv";
        sys.WriteContactInfoFile(f);
        std::ifstream in(f);
        std::string header, row, extra;
        ASSERT_TRUE((bool)std::getline(in, header));
        ASSERT_TRUE(header == "bi, bj, n_mag, fx, fy, fz");
        ASSERT_TRUE((bool)std::getline(in, row));
        ASSERT_TRUE(!std::getline(in, extra) || extra.empty());
        unsigned bi, bj;
        double nmag;
        char c;
        std::istringstream r(row);
        r >> bi >> c >> bj >> c >> nmag;
        ASSERT_TRUE(bi == 0 && bj == 1);
        ASSERT_NEAR(nmag, mass * g, 1e-3 * mass * g);
        // without recording the getters refuse, as upstream
        sys.SetRecordingContactInfo(false);
        bool threw = false;
        try { sys.getNormalForce(0, 1); } catch (const std::exception&) { threw = true; }
        ASSERT_TRUE(threw);
    }
    // ---- a sphere on a spinning floor
    {
        ChSystemDem sys(radius, density, ChVector3f(100.f, 100.f, 100.f));
        common_setup(sys, CHDEM_FRICTION_MODE::MULTI_STEP);
        const size_t plane = sys.CreateBCPlane(ChVector3f(0, 0, floor_z + 1.f), ChVector3f(0, 0, 1), false);
        sys.SetBCPlaneRotation(plane, ChVector3d(0, 0, floor_z + 1.f), ChVector3d(0, 0, 2.0));  // 2 rad/s about the z axis
        std::vector<ChVector3f> pts = {ChVector3f(5.f, 0, floor_z + 1.f + radius)};
        sys.SetParticles(pts);
        sys.Initialize();
        sys.AdvanceSimulation(0.3f);
        const ChVector3f v = sys.GetParticleVelocity(0);
        std::printf("sphere on the turntable: v = %g %g %g (surface speed there %g along +y)\n", v.x(), v.y(), v.z(), 2.0 * 5.0);
        ASSERT_TRUE(v.y() > 0.5);             // dragged along omega x r = (0, 10, 0)
        ASSERT_TRUE(std::fabs(v.x()) < 0.5 * v.y());
    }
    // ---- SINGLE_STEP checkpoint: partner map only
    {
        ChSystemDem sys(radius, density, ChVector3f(100.f, 100.f, 100.f));
        common_setup(sys, CHDEM_FRICTION_MODE::SINGLE_STEP);
        std::vector<ChVector3f> pts = {ChVector3f(0, 0, floor_z + radius), ChVector3f(0, 0, floor_z + 3 * radius)};
        sys.SetParticles(pts);
        sys.Initialize();
        sys.AdvanceSimulation(0.05f);
        const std::string cp = dir + "/single_step_checkpoint.dat";
        sys.WriteCheckpointFile(cp);
        std::ifstream in(cp);
        std::string line;
        bool saw = false;
        while (std::getline(in, line))
            if (line.find("partners") != std::string::npos) {
                saw = true;
                ASSERT_TRUE(line.find("history") == std::string::npos);
            }
        ASSERT_TRUE(saw);
        ChSystemDem again(cp);
        ASSERT_TRUE(again.GetNumParticles() == 2);
        // a file in the reference's layout, blank lines included (ReadHstHistory skips them, ChSystemDem.cpp:793-817)
        const std::string hf = dir + "/partners_only.hst";
        {
            std::ofstream o(hf);
            o << "\n  \npartners 12\n";
            for (int s = 0; s < 2; s++) {
                for (int k = 0; k < 12; k++) o << (k == 0 ? (s == 0 ? 1u : 0u) : (unsigned)NULL_CHDEM_ID) << " ";
                o << "\n\n";
            }
        }
        ChSystemDem third(radius, density, ChVector3f(100.f, 100.f, 100.f));
        common_setup(third, CHDEM_FRICTION_MODE::SINGLE_STEP);
        third.SetParticles(pts);
        third.ReadContactHistoryFile(hf);
        third.Initialize();
        third.AdvanceSimulation(0.01f);
        ASSERT_TRUE(third.GetParticlePosition(1).z() > floor_z);
    }
    // ---- setters after Initialize reach the engine (gravity, step size) or fail loudly (friction mode), never silently diverge
    {
        ChSystemDem sys(radius, density, ChVector3f(100.f, 100.f, 100.f));
        common_setup(sys, CHDEM_FRICTION_MODE::MULTI_STEP);
        std::vector<ChVector3f> pts = {ChVector3f(0, 0, 0)};
        sys.SetParticles(pts);
        sys.Initialize();
        sys.AdvanceSimulation(0.01f);
        const float vz1 = sys.GetParticleVelocity(0).z();
        ASSERT_NEAR(vz1, -980.f * 0.01f, 1e-3);
        sys.SetGravitationalAcceleration(ChVector3d(0, 0, 0));
        sys.AdvanceSimulation(0.01f);
        ASSERT_NEAR(sys.GetParticleVelocity(0).z(), vz1, 1e-6);  // free flight without gravity: the velocity stays
        sys.SetGravitationalAcceleration(ChVector3d(0, 0, -980));
        sys.SetFixedStepSize(5e-5f);
        const float t0 = sys.GetSimTime();
        sys.AdvanceSimulation(0.01f);  // 200 steps of the new size
        ASSERT_NEAR(sys.GetSimTime() - t0, 0.01, 1e-6);
        ASSERT_NEAR(sys.GetParticleVelocity(0).z(), vz1 - 980.f * 0.01f, 2e-3);
        bool threw = false;
        try { sys.SetFrictionMode(CHDEM_FRICTION_MODE::FRICTIONLESS); } catch (const std::exception&) { threw = true; }
        ASSERT_TRUE(threw);
    }
}
RUN_TEST(demContactInfo)

// API surface checks of the ChSystemDem mirror that the reference exercises through its demos rather than unit tests:
// particle-file formats (ChSystemDem_impl.cpp:212-333), checkpoint write -> read round trip (ChSystemDem.cpp:1322-1412,
// 704-857), queries (GetMaxParticleZ, GetNumParticleAboveZ, kinetic energy), moving boundary via SetBCOffsetFunction
// (demo_DEM_movingBoundary.cpp:119-121), cylinder container, DisableBCbyID, error behaviour.  argv[1] = scratch directory.
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "chrono_dem/physics/ChSystemDem.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

static void configure(ChSystemDem& s) {
    s.SetGravitationalAcceleration(ChVector3f(0, 0, -9.81f));
    s.SetFrictionMode(CHDEM_FRICTION_MODE::MULTI_STEP);
    s.SetTimeIntegrator(CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE);
    s.SetKn_SPH2SPH(1e5); s.SetKn_SPH2WALL(1e5);
    s.SetGn_SPH2SPH(5e2); s.SetGn_SPH2WALL(5e2);
    s.SetKt_SPH2SPH(3e4); s.SetKt_SPH2WALL(3e4);
    s.SetGt_SPH2SPH(50); s.SetGt_SPH2WALL(50);
    s.SetStaticFrictionCoeff_SPH2SPH(0.4f); s.SetStaticFrictionCoeff_SPH2WALL(0.4f);
    s.SetFixedStepSize(1e-4f);
    s.SetVerbosity(CHDEM_VERBOSITY::QUIET);
}

void demApi(int argc, char** argv) {
    ASSERT_TRUE(argc >= 2);
    const std::string dir = argv[1];
    const float R = 0.05f;
    std::vector<ChVector3f> pts;
    for (int k = 0; k < 4; k++)
        for (int j = 0; j < 6; j++)
            for (int i = 0; i < 6; i++)
                pts.push_back(ChVector3f(-0.3f + 0.11f * i + 0.003f * k, -0.3f + 0.11f * j, -0.45f + 0.11f * k));
    const size_t n = pts.size();

    // ---- settle a small bed, query, write files ----
    ChSystemDem a(R, 2500.f, ChVector3f(1.f, 1.f, 1.f));
    configure(a);
    a.SetParticles(pts);
    a.SetParticleOutputFlags(ABSV | VEL_COMPONENTS | FIXITY | ANG_VEL_COMPONENTS);
    size_t top = a.CreateBCPlane(ChVector3f(0, 0, 0.2f), ChVector3f(0, 0, -1), true);  // a lid, initially out of reach
    a.Initialize();
    ASSERT_TRUE(a.GetNumParticles() == n);
    double t = a.AdvanceSimulation(0.2f);
    ASSERT_NEAR(t, 0.2, 1e-6);
    ASSERT_NEAR(a.GetSimTime(), 0.2, 1e-5);
    ASSERT_TRUE(a.GetMaxParticleZ() < -0.45 + 0.11 * 3 + 1e-3);   // fell / compacted, nothing flew away
    ASSERT_TRUE(a.GetMinParticleZ() > -0.5 + 0.9 * R);             // floor holds
    ASSERT_TRUE(a.GetNumParticleAboveZ(-0.5f) == n);
    ASSERT_TRUE(a.GetNumParticleAboveZ(0.4f) == 0);
    ASSERT_TRUE(a.GetNumContacts() > n / 2);                       // stacked columns: ~3 sphere + 1 floor contact per column of 4
    ASSERT_TRUE(a.GetParticlesKineticEnergy() >= 0.f);

    a.WriteParticleFile(dir + "/p.csv");
    {
        std::ifstream f(dir + "/p.csv");
        std::string header;
        std::getline(f, header);
        ASSERT_TRUE(header == "x,y,z,vx,vy,vz,absv,fixed,wx,wy,wz");
        size_t rows = 0;
        std::string line;
        while (std::getline(f, line)) rows += !line.empty();
        ASSERT_TRUE(rows == n);
    }
    // force columns: (acceleration of the last step - g) * mass (ChSystemDem_impl.cpp:322-327), consistent with GetParticleLinAcc
    a.SetParticleOutputFlags(FORCE_COMPONENTS);
    a.WriteParticleFile(dir + "/pf.csv");
    a.SetParticleOutputFlags(ABSV | VEL_COMPONENTS | FIXITY | ANG_VEL_COMPONENTS);
    {
        std::ifstream f(dir + "/pf.csv");
        std::string header, line;
        std::getline(f, header);
        ASSERT_TRUE(header == "x,y,z,fx,fy,fz");
        const double mass = 4.0 / 3.0 * 3.14159265358979323846 * R * R * R * 2500.0;
        double fz_sum = 0, worst = 0;
        for (size_t i = 0; i < n; i++) {
            ASSERT_TRUE((bool)std::getline(f, line));
            for (char& ch : line) if (ch == ',') ch = ' ';
            std::istringstream is(line);
            double x, y, z, fx, fy, fz;
            is >> x >> y >> z >> fx >> fy >> fz;
            ASSERT_TRUE(!is.fail());
            const ChVector3f acc = a.GetParticleLinAcc((int)i);
            const double want[3] = {acc.x() * mass, acc.y() * mass, (acc.z() + 9.81f) * mass};
            const double got[3] = {fx, fy, fz};
            for (int k = 0; k < 3; k++)
                worst = std::max(worst, std::abs(got[k] - want[k]) / (mass * 9.81));
            fz_sum += fz;
        }
        std::printf("force columns vs GetParticleLinAcc: worst difference %.3g of a sphere's weight; sum fz / bed weight %.4f\n",
                    worst, fz_sum / (n * mass * 9.81));
        ASSERT_TRUE(worst < 1e-4);                                     // float getter + 6-digit text
        ASSERT_NEAR(fz_sum / (n * mass * 9.81), 1.0, 0.25);            // a bed (nearly) at rest carries its weight
    }
    // contact partners of a bottom-layer sphere of the settled columns: the sphere above it and the floor (BC 4 = bottom
    // z plane, label nSpheres + BC_id + 1: ChDemBoundaryConditions.cuh:102)
    {
        std::vector<unsigned int> nb;
        a.getNeighbors(0, nb);
        bool has_above = false, has_wall = false;
        for (unsigned int l : nb) {
            has_above |= (l == 36u);
            has_wall |= (l > n);
        }
        std::printf("sphere 0 touches %zu partners\n", nb.size());
        ASSERT_TRUE(has_above && has_wall);
    }
    a.SetParticleOutputMode(CHDEM_OUTPUT_MODE::BINARY);
    a.WriteParticleFile(dir + "/p.raw");
    {
        std::ifstream f(dir + "/p.raw", std::ios::binary | std::ios::ate);
        ASSERT_TRUE((size_t)f.tellg() == n * (3 + 3 + 1 + 3) * sizeof(float));
    }

    // ---- moving boundary: push the lid down onto the bed and read the reaction force ----
    ChVector3f f0;
    ASSERT_TRUE(a.GetBCReactionForces(top, f0));
    ASSERT_NEAR(f0.z(), 0.0, 1e-12);
    const float speed = 0.5f;
    a.SetBCOffsetFunction(top, [speed](float tt) { return make_double3(0, 0, -(double)speed * (tt - 0.2)); });
    ChVector3f f1;
    int chunks = 0;
    for (; chunks < 150; chunks++) {  // the bed keeps compacting while the lid comes down: advance until they meet
        a.AdvanceSimulation(0.01f);
        ASSERT_TRUE(a.GetBCReactionForces(top, f1));
        if (f1.z() > 0.0f)
            break;
    }
    std::printf("lid touched the bed after %d chunks: lid z %.5f, top sphere z %.5f, reaction %g %g %g\n", chunks,
                a.GetBCPlanePosition(top).z(), a.GetMaxParticleZ(), f1.x(), f1.y(), f1.z());
    ASSERT_TRUE(f1.z() > 0.0f);  // spheres push the lid upwards
    ASSERT_TRUE(a.GetBCPlanePosition(top).z() - a.GetMaxParticleZ() < R);
    ASSERT_NEAR(a.GetBCPlanePosition(top).z(), 0.2 - speed * (a.GetSimTime() - 0.2), 1e-4);
    ASSERT_TRUE(a.DisableBCbyID(top));
    ASSERT_TRUE(!a.DisableBCbyID(1000));

    // ---- checkpoint round trip: b restarts from a's checkpoint and follows the same trajectory ----
    a.WriteCheckpointFile(dir + "/cp.dat");
    ChSystemDem b(dir + "/cp.dat");
    b.SetVerbosity(CHDEM_VERBOSITY::QUIET);
    size_t top_b = b.CreateBCPlane(ChVector3f(0, 0, 0.2f), ChVector3f(0, 0, -1), true);
    b.DisableBCbyID(top_b);
    b.Initialize();
    ASSERT_TRUE(b.GetNumParticles() == n);
    ASSERT_NEAR(b.GetSimTime(), a.GetSimTime(), 1e-5);
    ASSERT_TRUE(b.GetNumContacts() == 0);          // history is staged, contacts appear with the first step
    a.AdvanceSimulation(0.01f);
    b.AdvanceSimulation(0.01f);
    ASSERT_TRUE(b.GetNumContacts() > n / 2);
    double dmax = 0;
    for (size_t i = 0; i < n; i++)
        dmax = std::max(dmax, (double)(a.GetParticlePosition((int)i) - b.GetParticlePosition((int)i)).Length());
    std::printf("checkpoint restart: max position difference after 100 steps %.3g\n", dmax);
    ASSERT_TRUE(dmax < 2e-5);  // the text checkpoint keeps 6 significant digits

    // ---- cylinder container + fixed particle ----
    ChSystemDem c(R, 2500.f, ChVector3f(2.f, 2.f, 1.f));
    configure(c);
    std::vector<ChVector3f> one = {ChVector3f(0.25f, 0, -0.4f), ChVector3f(0, 0, 0.3f)};
    std::vector<ChVector3f> v1 = {ChVector3f(1.0f, 0, 0), ChVector3f(0, 0, 0)};
    c.SetParticles(one, v1);
    c.SetParticleFixed({false, true});
    c.CreateBCCylinderZ(ChVector3f(0, 0, 0), 0.4f, false, false);
    c.Initialize();
    ASSERT_NEAR(c.GetParticleLinAcc(0).Length(), 0.0, 0.0);      // no step yet
    c.AdvanceSimulation(1e-4f);
    ASSERT_NEAR(c.GetParticleLinAcc(0).z(), -9.81, 1e-5);         // free flight: gravity alone
    ASSERT_NEAR(c.GetParticleLinAcc(0).x(), 0.0, 1e-6);
    ASSERT_NEAR(c.GetParticleLinAcc(1).Length(), 0.0, 0.0);      // fixed
    c.AdvanceSimulation(0.5f);
    ChVector3f p0 = c.GetParticlePosition(0), p1 = c.GetParticlePosition(1);
    ASSERT_TRUE(std::sqrt(p0.x() * p0.x() + p0.y() * p0.y()) < 0.4 - R + 0.01);  // bounced off the cylinder wall
    ASSERT_NEAR(p1.z(), 0.3, 1e-6);                                               // the fixed sphere did not move
    ASSERT_TRUE(c.IsFixed(1) && !c.IsFixed(0));

    // ---- error behaviour: wrong call order must not pass silently ----
    bool threw = false;
    try {
        ChSystemDem d(R, 2500.f, ChVector3f(1, 1, 1));
        d.AdvanceSimulation(0.1f);
    } catch (const std::exception&) {
        threw = true;
    }
    ASSERT_TRUE(threw);
}
RUN_TEST(demApi)

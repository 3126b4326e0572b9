// CPU-only checks of the look-alike utilities the Chrono::Dem demos rely on (no engine call): point samplers
// (include/chrono/utils/ChUtilsSamplers.h), JSON parameter reader (include/chrono_dem/utils/ChDemJsonParser.h), data-path
// helpers, OBJ reading of ChTriangleMeshConnected.
#include <algorithm>
#include <fstream>
#include <string>
#include <vector>
#include "chrono/core/ChDataPath.h"
#include "chrono/geometry/ChTriangleMeshConnected.h"
#include "chrono/utils/ChUtilsSamplers.h"
#include "chrono_dem/utils/ChDemJsonParser.h"
#include "mini_test.h"

using namespace chrono;
using namespace chrono::dem;

template <class T>
static double min_distance(const std::vector<ChVector3<T>>& p) {
    double m = 1e300;
    for (size_t i = 0; i < p.size(); i++)
        for (size_t j = i + 1; j < p.size(); j++)
            m = std::min(m, (double)(p[i] - p[j]).Length());
    return m;
}

void utils_check(int argc, char** argv) {
    ASSERT_TRUE(argc > 2);
    const std::string dir = argv[1], json = argv[2];

    // ---- HCP: nearest-neighbour distance is the separation; 12 neighbours in the bulk
    utils::ChHCPSampler<double> hcp(0.1);
    auto pts = hcp.SampleBox(ChVector3d(0, 0, 0), ChVector3d(0.5, 0.5, 0.5));
    ASSERT_TRUE(pts.size() > 1200 && pts.size() < 1700);
    ASSERT_NEAR(min_distance(pts), 0.1, 1e-9);
    int best = 0;
    for (auto& c : pts) {
        if (std::max({std::abs(c.x()), std::abs(c.y()), std::abs(c.z())}) > 0.25) continue;
        int nb = 0;
        for (auto& q : pts) nb += ((c - q).Length() < 0.1001 && (c - q).Length() > 1e-9);
        best = std::max(best, nb);
    }
    ASSERT_TRUE(best == 12);
    for (auto& p : hcp.SampleCylinderZ(ChVector3d(1, 2, 3), 0.4, 0.0))
        ASSERT_TRUE(std::hypot(p.x() - 1, p.y() - 2) <= 0.4 + 1e-6 && std::abs(p.z() - 3) < 1e-6);

    // ---- grid
    utils::ChGridSampler<float> grid(0.25f);
    ASSERT_TRUE(grid.SampleBox(ChVector3f(0, 0, 0), ChVector3f(0.5f, 0.5f, 0.5f)).size() == 125);

    // ---- Poisson disk: minimum distance respected, volume filled to a sensible density, reproducible
    utils::ChPDSampler<float> pd(0.1f);
    auto a = pd.SampleBox(ChVector3f(0, 0, 0), ChVector3f(0.5f, 0.5f, 0.25f));
    ASSERT_TRUE(min_distance(a) >= 0.1 - 1e-5);
    ASSERT_TRUE(a.size() > 280 && a.size() < 450);  // volume 0.5 = 500 s^3: Bridson sampling saturates near 0.7 points per s^3
    utils::ChPDSampler<float> pd2(0.1f);
    ASSERT_TRUE(pd2.SampleBox(ChVector3f(0, 0, 0), ChVector3f(0.5f, 0.5f, 0.25f)).size() == a.size());
    auto layers = utils::ChPDLayerSamplerBox<float>(ChVector3f(0, 0, 1), ChVector3f(1, 1, 0.25f), 0.2f, 1.05f);
    ASSERT_TRUE(layers.size() > 150);
    ASSERT_TRUE(min_distance(layers) >= 0.2 * 1.05 - 1e-5);
    for (auto& p : layers) ASSERT_TRUE(p.z() >= 0.75f - 1e-5f && p.z() < 1.25f);

    // ---- JSON parameter file (fixture re-typed from data/dem/mixer.json)
    ChDemSimulationParameters prm{};
    prm.box_Y = -1.f;
    ASSERT_TRUE(ParseJSON(json, prm, false));
    ASSERT_NEAR(prm.sphere_radius, 2.0, 1e-6);
    ASSERT_NEAR(prm.normalStiffS2M, 1e8, 1e-3);
    ASSERT_NEAR(prm.step_size, 1e-5, 1e-12);
    ASSERT_TRUE(prm.psi_T == 32 && prm.psi_L == 16);
    ASSERT_TRUE(prm.output_dir == "mixer" && prm.write_mode == CHDEM_OUTPUT_MODE::CSV);
    ASSERT_NEAR(prm.box_Y, -1.0, 0);  // absent key keeps its value
    {
        std::ofstream bad(dir + "/bad.json");
        bad << "{ \"write_mode\": \"parquet\" }";
    }
    ASSERT_TRUE(!ParseJSON(dir + "/bad.json", prm, false));
    {
        std::ofstream nested(dir + "/nested.json");
        nested << "{ \"a\": [1, 2, {\"b\": \"x\"}], \"time_end\": 2.5e0, \"verbose\": false, \"c\": {\"d\": null} }";
    }
    ASSERT_TRUE(ParseJSON(dir + "/nested.json", prm, false));
    ASSERT_NEAR(prm.time_end, 2.5, 1e-6);
    ASSERT_TRUE(!ParseJSON(dir + "/does_not_exist.json", prm, false));

    // ---- data path + OBJ
    SetChronoDataPath(dir + "/");
    ASSERT_TRUE(GetChronoDataFile("m.obj") == dir + "/m.obj");
    ASSERT_TRUE(CreateOutputDirectory(dir + "/out/deeper"));
    {
        std::ofstream obj(GetChronoDataFile("m.obj"));
        obj << "# quad + triangle\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n";
    }
    ChTriangleMeshConnected mesh;
    ASSERT_TRUE(mesh.LoadWavefrontMesh(GetChronoDataFile("m.obj")));
    ASSERT_TRUE(mesh.GetNumTriangles() == 3 && mesh.GetNumVertices() == 4);
    mesh.Transform(ChVector3d(0, 0, 5), ChMatrix33<double>(2.0));
    ASSERT_NEAR(mesh.GetTriangle(0).p3.x(), 2.0, 1e-12);
    ASSERT_NEAR(mesh.GetTriangle(2).p1.z(), 5.0, 1e-12);
}
RUN_TEST(utils_check)

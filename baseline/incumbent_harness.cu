// =============================================================================
// incumbent_harness.cu -- BENCHMARK INFRASTRUCTURE (not product code, never linked into libchrono_b200_dem.so).
//
// Drives the UNMODIFIED Chrono::Dem CUDA implementation -- /root/reference/src/chrono_dem/gpu/ChDemSMC.cu (kernels +
// ChSystemDem_impl::AdvanceSimulation, :619-691) and physics/ChSystemDem_impl.cpp (initializeSpheres, :1130-1166),
// compiled where they lie for sm_100a by baseline/Makefile -- on the workload of BASELINE.json configs[1], so that
// bench.py can report "the existing kernels on the same B200" (`incumbent_gpu`) next to ours.
//
// The reference's public front end (ChSystemDem.cpp) needs Chrono core (Eigen3: absent here), so the harness plays its
// role: ChSystemDem is a friend of ChSystemDem_impl (ChSystemDem_impl.h:790) and every setter of the front end is a
// one-line store into the impl (ChSystemDem.cpp:52-260); the class below does exactly those stores and nothing else.
//
// Scene = the generator of chrono_b200/scenes.py::settling_scene (jittered HCP lattice at spacing 2R in an open box)
// restated here so that the binary needs no input file: same lattice, same spacing, same material; the jitter comes
// from a different generator (the incumbent works in int32/fp32 lattice units: bit-identical inputs are moot).
//
// Output: one JSON line on stdout.
// =============================================================================
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "chrono_dem/physics/ChSystemDem_impl.h"

namespace chrono {
namespace dem {

// Stand-in for the reference's front-end class: same name, so that the friend declaration of ChSystemDem_impl applies.
class ChSystemDem {
  public:
    ChSystemDem(float rad, float density, float3 box, float3 O) : m_sys(new ChSystemDem_impl(rad, density, box, O)) {}
    ~ChSystemDem() { delete m_sys; }
    ChSystemDem_impl* m_sys;

    void Configure(float dt, double young, double poisson, double cor, float mu, bool frictionless) {
        ChSystemDem_impl* s = m_sys;
        s->verbosity = CHDEM_VERBOSITY::QUIET;
        s->X_accGrav = 0.f; s->Y_accGrav = 0.f; s->Z_accGrav = -9.81f;                 // SetGravitationalAcceleration
        s->stepSize_UU = dt;                                                              // SetFixedStepSize
        s->BD_is_fixed = true;                                                            // SetBDFixed
        s->use_mat_based = true; s->gran_params->use_mat_based = true;                    // UseMaterialBasedModel
        s->YoungsModulus_sphere_UU = young; s->YoungsModulus_wall_UU = young;             // SetYoungModulus_*
        s->PoissonRatio_sphere_UU = poisson; s->PoissonRatio_wall_UU = poisson;           // SetPoissonRatio_*
        s->COR_sphere_UU = cor; s->COR_wall_UU = cor;                                     // SetRestitution_*
        s->gran_params->static_friction_coeff_s2s = mu;                                   // SetStaticFrictionCoeff_*
        s->gran_params->static_friction_coeff_s2w = mu;
        s->gran_params->friction_mode = frictionless ? CHDEM_FRICTION_MODE::FRICTIONLESS : CHDEM_FRICTION_MODE::MULTI_STEP;
        s->gran_params->rolling_mode = CHDEM_ROLLING_MODE::NO_RESISTANCE;
        s->gran_params->time_integrator = CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE;     // SetTimeIntegrator
        s->time_integrator = CHDEM_TIME_INTEGRATOR::CENTERED_DIFFERENCE;
        s->cohesion_over_gravity = 0.f; s->adhesion_s2w_over_gravity = 0.f;
        s->gran_params->recording_contactInfo = false;
    }
    void SetParticles(const std::vector<float3>& p) { m_sys->SetParticles(p); }
    void Initialize() { m_sys->initializeSpheres(); }
    double Advance(float duration) { return m_sys->AdvanceSimulation(duration); }
    unsigned NumContacts() const { return m_sys->GetNumContacts(); }
    float3 Position(int i) const { return m_sys->GetParticlePosition(i); }
    double MaxZ() { return m_sys->GetMaxParticleZ(true); }
};

}  // namespace dem
}  // namespace chrono

// ChHCPSampler lattice in generation order (z layers, y rows, x), cf. chrono_b200/scenes.py::hcp_points
static void hcp_points(const double lo[3], const double hi[3], double sep, size_t want, std::vector<float3>& out,
                       double jitter, unsigned seed) {
    const double dx = sep, dy = sep * (std::sqrt(3.0) / 2), dz = sep * std::sqrt(2.0 / 3.0);
    const int nx = (int)((hi[0] - lo[0]) / dx) + 1, ny = (int)((hi[1] - lo[1]) / dy) + 1, nz = (int)((hi[2] - lo[2]) / dz) + 1;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-jitter, jitter);
    for (int k = 0; k < nz && out.size() < want; k++)
        for (int j = 0; j < ny && out.size() < want; j++)
            for (int i = 0; i < nx && out.size() < want; i++) {
                const double offy = (k % 2 == 0) ? 0.0 : dy / 3, offx = ((j + k) % 2 == 0) ? 0.0 : dx / 2;
                const double x = lo[0] + offx + i * dx, y = lo[1] + offy + j * dy, z = lo[2] + k * dz;
                if (x > hi[0] + 1e-12 || y > hi[1] + 1e-12 || z > hi[2] + 1e-12)
                    continue;
                out.push_back(make_float3((float)(x + U(rng)), (float)(y + U(rng)), (float)(z + U(rng))));
            }
}

int main(int argc, char** argv) {
    size_t n = 1000000;
    int warm = 100, steps = 300;
    bool frictionless = false;
    double sep_factor = 2.0;
    for (int a = 1; a < argc; a++) {
        if (!strcmp(argv[a], "--spheres") && a + 1 < argc) n = (size_t)atoll(argv[++a]);
        else if (!strcmp(argv[a], "--warmup") && a + 1 < argc) warm = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--steps") && a + 1 < argc) steps = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--sep") && a + 1 < argc) sep_factor = atof(argv[++a]);
        else if (!strcmp(argv[a], "--frictionless")) frictionless = true;
    }
    const double R = 0.02, rho = 2000.0, dt = 1e-4;
    const double sep = sep_factor * R;
    const double site = sep * sep * sep / std::sqrt(2.0);
    const double L = std::cbrt((double)n * site / 0.26);
    const int per_layer = ((int)((L - 2 * R) / sep) + 1) * ((int)((L - 2 * R) / (sep * std::sqrt(3.0) / 2)) + 1);
    const int layers = (int)std::ceil((double)n / per_layer) + 2;
    const double lo[3] = {-L / 2 + R * 1.01, -L / 2 + R * 1.01, R * 1.01};
    const double hi[3] = {L / 2 - R * 1.01, L / 2 - R * 1.01, lo[2] + layers * sep * std::sqrt(2.0 / 3.0)};
    std::vector<float3> pts;
    pts.reserve(n);
    hcp_points(lo, hi, sep, n, pts, 0.005 * R, 12346u);
    if (pts.size() < n) {
        printf("{\"impl\": \"incumbent\", \"unavailable\": \"lattice smaller than requested\"}\n");
        return 0;
    }
    double zmax = 0;
    for (auto& p : pts)
        zmax = std::max(zmax, (double)p.z);
    const double Lz = std::max(zmax + 2 * R, 0.25 * L) * 1.25;
    // Chrono::Dem's big domain is a closed box centred at O (ChSystemDem_impl.cpp:102-132): put its floor at z = 0
    chrono::dem::ChSystemDem sys((float)R, (float)rho, make_float3((float)L, (float)L, (float)Lz), make_float3(0.f, 0.f, (float)(Lz / 2)));
    sys.Configure((float)dt, 2e6, 0.3, 0.4, 0.4f, frictionless);
    sys.SetParticles(pts);
    sys.Initialize();
    sys.Advance((float)(warm * dt));
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0);
    sys.Advance((float)(steps * dt));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    auto t1 = std::chrono::steady_clock::now();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double wall = std::chrono::duration<double>(t1 - t0).count();
    const unsigned nc = frictionless ? 0u : sys.NumContacts();
    cudaError_t err = cudaGetLastError();
    printf("{\"impl\": \"incumbent\", \"what\": \"unmodified Chrono::Dem CUDA (ChDemSMC.cu, sm_100a build), %s, int32/fp32 arithmetic\", "
           "\"spheres\": %zu, \"timesteps\": %d, \"warmup_timesteps\": %d, \"ms_device\": %.3f, \"s_wall\": %.4f, "
           "\"value\": %.1f, \"unit\": \"sphere-steps/s\", \"contacts_per_sphere\": %.3f, \"max_z\": %.4f, \"cuda_error\": \"%s\"}\n",
           frictionless ? "frictionless material-based" : "MULTI_STEP friction, material-based Hertz", n, steps, warm, ms, wall,
           (double)n * steps / wall, 2.0 * nc / (double)n, sys.MaxZ(), err == cudaSuccess ? "" : cudaGetErrorString(err));
    return 0;
}

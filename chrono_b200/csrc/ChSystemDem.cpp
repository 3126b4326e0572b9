// =============================================================================
// ChSystemDem.cpp -- host-side mirror of chrono::dem::ChSystemDem (reference: src/chrono_dem/physics/ChSystemDem.cpp,
// ChSystemDem_impl.cpp) written on top of the C ABI in include/chrono_b200_dem.h.  No CUDA in this file: everything
// that touches the device goes through dem_b200_* entry points.
// =============================================================================
#include "chrono_dem/physics/ChSystemDem.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "chrono_b200_dem.h"

namespace chrono {
namespace dem {

namespace {
constexpr double kPi = 3.141592653589793238462643383279;

[[noreturn]] void fail(const std::string& msg) {
    // the reference prints and calls exit(1) (CHDEM_ERROR, src/chrono_dem/utils/ChDemUtilities.h:20-26)
    throw std::runtime_error("ChSystemDem: " + msg);
}
}  // namespace

enum class BCKind { PLANE, ZCYL, SPHERE, ZCONE, PLATE };

struct BCInfo {
    BCKind kind = BCKind::PLANE;
    double pos[3] = {0, 0, 0};     // reference position (plane point / cylinder axis point), user units
    double normal[3] = {0, 0, 1};  // plane normal
    double radius = 0;             // cylinder / ball
    double slope = 0, hmin = 0, hmax = 0;  // cone
    double vel[3] = {0, 0, 0};     // ball velocity (SetBCSphereVelocity)
    double mass = 0;               // ball: > 0 -> the ball is a free body driven by the spheres' reaction force and gravity
    double rot_center[3] = {0, 0, 0}, rot_omega[3] = {0, 0, 0};  // plane: SetBCPlaneRotation
    bool has_rotation = false;
    bool spheres_inside = true;
    bool track_forces = false;
    bool enabled = true;
    bool has_offset = false;
    GranPositionFunction offset = GranPosFunction_default;
    int wall = -1;  // engine wall index
};

class ChSystemDem_impl {
  public:
    dem_b200_system* h = nullptr;
    bool initialized = false;
    // geometry / material (user units)
    float radius = 0, density = 0;
    float box[3] = {0, 0, 0};
    float O[3] = {0, 0, 0};
    bool BD_fixed = true;
    float step = 1e-4f;
    float grav[3] = {0, 0, 0};
    double elapsed = 0;
    CHDEM_VERBOSITY verbosity = CHDEM_VERBOSITY::INFO;
    CHDEM_OUTPUT_MODE out_mode = CHDEM_OUTPUT_MODE::CSV;
    unsigned int out_flags = ABSV;
    CHDEM_TIME_INTEGRATOR integrator = CHDEM_TIME_INTEGRATOR::EXTENDED_TAYLOR;
    CHDEM_FRICTION_MODE friction = CHDEM_FRICTION_MODE::FRICTIONLESS;
    CHDEM_ROLLING_MODE rolling = CHDEM_ROLLING_MODE::NO_RESISTANCE;
    bool use_mat_based = false;
    bool use_min_length = true, defragment = true, record_contacts = false;
    unsigned psi_T = 32, psi_L = 16;
    float psi_R = 1.f, max_safe_vel = (float)UINT_MAX;
    // per-class user coefficients: [0] sphere-sphere, [1] sphere-wall, [2] sphere-mesh
    double Kn[3] = {0, 0, 0}, Kt[3] = {0, 0, 0}, Gn[3] = {0, 0, 0}, Gt[3] = {0, 0, 0};
    double mu[3] = {0, 0, 0}, mu_roll[3] = {0, 0, 0}, mu_spin[3] = {0, 0, 0};
    double cohesion_over_g = 0, adhesion_over_g[3] = {0, 0, 0};
    double young[3] = {0, 0, 0}, poisson[3] = {0, 0, 0}, cor[3] = {0, 0, 0};
    // particles (staged until Initialize)
    std::vector<double> pos, vel, omg, rad;
    std::vector<uint8_t> fixed;
    struct HistRow { uint32_t sphere, partner; bool is_bc; double d[3]; };
    std::vector<HistRow> hist;  // staged contact history (checkpoint): sphere index, partner sphere index or BC id
    std::vector<BCInfo> bcs;
    bool any_offset = false;
    // triangle meshes staged by ChSystemDemMesh (uploaded at Initialize, before the spheres: their triangles take the
    // shape ids between the walls and the spheres)
    struct MeshStage {
        std::vector<double> verts9;  // 9 doubles per triangle, mesh frame
        double mass = 1.0;
        bool has_motion = false;
        double pos[3] = {0, 0, 0}, rot[4] = {1, 0, 0, 0}, lin[3] = {0, 0, 0}, ang[3] = {0, 0, 0};
    };
    std::vector<MeshStage> meshes;
    bool mesh_collision = true;
    uint32_t num_triangles() const {
        size_t t = 0;
        for (auto& m : meshes) t += m.verts9.size() / 9;
        return (uint32_t)t;
    }

    size_t n() const { return rad.empty() ? pos.size() / 3 : rad.size(); }

    void check(int rc, const char* what) const {
        if (rc < 0)
            fail(std::string(what) + ": " + dem_b200_last_error(h));
    }
    double sphere_mass() const { return 4.0 / 3.0 * kPi * (double)radius * radius * radius * density; }
    double gmag() const { return std::sqrt((double)grav[0] * grav[0] + (double)grav[1] * grav[1] + (double)grav[2] * grav[2]); }

    // Dem F = K d^{3/2} R^{-1/2}  <->  engine (Multicore Hertz, user coefficients)
    // F = kn R* d^{3/2}; R* = R/2 sphere-sphere, R sphere-wall / sphere-mesh.
    dem_b200_contact_class make_class(int c) const {
        dem_b200_contact_class k{};
        const double R = radius;
        const double f = ((c == 0) ? 2.0 : 1.0) / std::pow(R, 1.5);
        k.kn = f * Kn[c]; k.kt = f * Kt[c]; k.gn = f * Gn[c]; k.gt = f * Gt[c];
        const bool fr = friction != CHDEM_FRICTION_MODE::FRICTIONLESS;
        k.mu = fr ? mu[c] : 0.0;
        const bool roll = fr && rolling == CHDEM_ROLLING_MODE::SCHWARTZ;
        k.mu_roll = roll ? mu_roll[c] : 0.0;
        k.mu_spin = 0.0;  // accepted but unused by the reference
        k.cr = std::min(cor[0], cor[c == 0 ? 0 : c]);
        // constant cohesion: ratio * m * |g| (ChSystemDem_impl.cpp:1444-1445)
        k.adhesion = ((c == 0) ? cohesion_over_g : adhesion_over_g[c]) * sphere_mass() * gmag();
        // material-based model (utils/ChDemUtilities.h:65-70)
        const int w = (c == 0) ? 0 : c;
        if (young[0] > 0 && young[w] > 0) {
            const double invE = (1 - poisson[0] * poisson[0]) / young[0] + (1 - poisson[w] * poisson[w]) / young[w];
            const double invG = 2 * (2 - poisson[0]) * (1 + poisson[0]) / young[0] + 2 * (2 - poisson[w]) * (1 + poisson[w]) / young[w];
            k.E_eff = 1 / invE;
            k.G_eff = 1 / invG;
        }
        return k;
    }

    dem_b200_config make_config() const {
        dem_b200_config c{};
        c.device = 0;
        c.force_model = DEMB200_HERTZ;
        c.adhesion_model = DEMB200_ADH_CONSTANT;
        c.tangential_mode = friction == CHDEM_FRICTION_MODE::MULTI_STEP ? DEMB200_TANG_MULTISTEP
                            : friction == CHDEM_FRICTION_MODE::SINGLE_STEP ? DEMB200_TANG_ONESTEP : DEMB200_TANG_NONE;
        c.use_mat_props = use_mat_based ? 1 : 0;
        c.integrator = (int)integrator;  // same order as CHDEM_TIME_INTEGRATOR
        c.history_slots = 16;
        c.char_vel = 1.0; c.min_slip_vel = 1e-4; c.min_roll_vel = 1e-4; c.min_spin_vel = 1e-4;
        c.dt = step;
        for (int k = 0; k < 3; k++) c.gravity[k] = grav[k];
        // Multicore bins only matter for the parity inspection calls
        c.bins_per_axis[0] = c.bins_per_axis[1] = c.bins_per_axis[2] = 10;
        for (int k = 0; k < 3; k++) {
            dem_b200_material& m = c.material[k];
            m.young = (float)std::max(young[k], 1.0); m.poisson = (float)poisson[k];
            m.mu_s = (float)mu[k]; m.cr = (float)cor[k];
        }
        c.mass_coef = 4.0 / 3.0 * kPi * density;
        c.wall_mass = 1e30;  // Dem walls are infinitely heavy: m_eff = m (ChDemBoundaryConditions.cuh:446)
        c.mesh_mass = 1e30;
        c.verlet_skin = -1.0;
        // candidate slots per sphere hold sphere candidates AND the mesh facets in reach: leave room for fine meshes
        c.neighbor_slots = meshes.empty() ? 0 : 48;
        return c;
    }

    // A setter called after Initialize: the reference stores into GranParams, which its kernels read every step
    // (ChSystemDem.cpp:52-260); here the change is pushed through the ABI.  What the engine cannot change on a running system
    // (friction mode, force model, to / from CHUNG, the slot counts) comes back as an error instead of being ignored.
    void refresh(const char* what) {
        if (!initialized)
            return;
        dem_b200_config cfg = make_config();
        check(dem_b200_set_config(h, &cfg), what);
        for (int k = 0; k < 3; k++) {
            dem_b200_contact_class cc = make_class(k);
            check(dem_b200_set_contact_class(h, k, &cc), what);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------

ChSystemDem::ChSystemDem(float sphere_rad, float density, const ChVector3f& boxDims, ChVector3f O) : m_RTF(0) {
    m_sys = new ChSystemDem_impl();
    m_sys->radius = sphere_rad;
    m_sys->density = density;
    m_sys->box[0] = boxDims.x(); m_sys->box[1] = boxDims.y(); m_sys->box[2] = boxDims.z();
    m_sys->O[0] = O.x(); m_sys->O[1] = O.y(); m_sys->O[2] = O.z();
    m_sys->bcs.resize(NUM_RESERVED_BC_IDS);  // big-domain walls, created at Initialize (ChSystemDem_impl.cpp:102-132)
}

ChSystemDem::ChSystemDem(const std::string& checkpoint) : m_RTF(0) {
    m_sys = new ChSystemDem_impl();
    m_sys->bcs.resize(NUM_RESERVED_BC_IDS);
    ReadCheckpointFile(checkpoint, true);
}

ChSystemDem::~ChSystemDem() {
    if (m_sys) {
        if (m_sys->h)
            dem_b200_destroy(m_sys->h);
        delete m_sys;
    }
}

void* ChSystemDem::GetEngineHandle() const { return m_sys->h; }

// ---- setters ---------------------------------------------------------------------------------------------------------
void ChSystemDem::SetGravitationalAcceleration(const ChVector3f& g) {
    m_sys->grav[0] = g.x(); m_sys->grav[1] = g.y(); m_sys->grav[2] = g.z();
    m_sys->refresh("SetGravitationalAcceleration");
}

void ChSystemDem::SetParticles(const std::vector<ChVector3f>& points, const std::vector<ChVector3f>& vels,
                               const std::vector<ChVector3f>& ang_vels) {
    if (m_sys->initialized)
        fail("SetParticles after Initialize");
    if (!vels.empty() && vels.size() != points.size())
        fail("SetParticles: velocities and positions differ in size");
    if (!ang_vels.empty() && ang_vels.size() != points.size())
        fail("SetParticles: angular velocities and positions differ in size");
    const size_t n = points.size();
    m_sys->pos.resize(3 * n);
    m_sys->vel.assign(3 * n, 0.0);
    m_sys->omg.assign(3 * n, 0.0);
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            m_sys->pos[3 * i + k] = points[i][k];
            if (!vels.empty()) m_sys->vel[3 * i + k] = vels[i][k];
            if (!ang_vels.empty()) m_sys->omg[3 * i + k] = ang_vels[i][k];
        }
    if (m_sys->rad.size() != n)
        m_sys->rad.assign(n, m_sys->radius);
    if (m_sys->fixed.size() != n)
        m_sys->fixed.assign(n, 0);
}

void ChSystemDem::SetParticleRadii(const std::vector<float>& radii) {
    if (m_sys->initialized)
        fail("SetParticleRadii after Initialize");
    m_sys->rad.assign(radii.begin(), radii.end());
}

void ChSystemDem::SetBDFixed(bool fixed) { m_sys->BD_fixed = fixed; }
void ChSystemDem::SetBDCenter(const ChVector3f& O) { m_sys->O[0] = O.x(); m_sys->O[1] = O.y(); m_sys->O[2] = O.z(); }
void ChSystemDem::SetParticleFixed(const std::vector<bool>& fixed) {
    m_sys->fixed.resize(fixed.size());
    for (size_t i = 0; i < fixed.size(); i++) m_sys->fixed[i] = fixed[i] ? 1 : 0;
}
void ChSystemDem::SetParticleOutputMode(CHDEM_OUTPUT_MODE mode) { m_sys->out_mode = mode; }
void ChSystemDem::SetParticleOutputFlags(unsigned int flags) { m_sys->out_flags = flags; }
void ChSystemDem::SetFixedStepSize(float size_UU) { m_sys->step = size_UU; m_sys->refresh("SetFixedStepSize"); }
float ChSystemDem::GetFixedStepSize() const { return m_sys->step; }
void ChSystemDem::SetDefragmentOnInitialize(bool defragment) { m_sys->defragment = defragment; }
void ChSystemDem::EnableMinLength(bool useMinLen) { m_sys->use_min_length = useMinLen; }
void ChSystemDem::SetTimeIntegrator(CHDEM_TIME_INTEGRATOR new_integrator) {
    const auto old = m_sys->integrator;
    m_sys->integrator = new_integrator;
    try {
        m_sys->refresh("SetTimeIntegrator");
    } catch (...) {
        m_sys->integrator = old;  // the engine refused (not changeable on a running system): keep the mirror consistent with it
        throw;
    }
}
void ChSystemDem::SetFrictionMode(CHDEM_FRICTION_MODE new_mode) {
    const auto old = m_sys->friction;
    m_sys->friction = new_mode;
    try {
        m_sys->refresh("SetFrictionMode");
    } catch (...) {
        m_sys->friction = old;  // the engine refused (not changeable on a running system): keep the mirror consistent with it
        throw;
    }
}
void ChSystemDem::SetRollingMode(CHDEM_ROLLING_MODE new_mode) {
    if (new_mode == CHDEM_ROLLING_MODE::ELASTIC_PLASTIC)
        fail("ELASTIC_PLASTIC rolling is not implemented (nor in the reference, ChDemHelpers.cuh:236-240)");
    m_sys->rolling = new_mode;
    m_sys->refresh("SetRollingMode");
}
void ChSystemDem::SetStaticFrictionCoeff_SPH2SPH(float mu) { m_sys->mu[0] = mu; m_sys->refresh("SetStaticFrictionCoeff_SPH2SPH"); }
void ChSystemDem::SetStaticFrictionCoeff_SPH2WALL(float mu) { m_sys->mu[1] = mu; m_sys->refresh("SetStaticFrictionCoeff_SPH2WALL"); }
void ChSystemDem::SetRollingCoeff_SPH2SPH(float mu) { m_sys->mu_roll[0] = mu; m_sys->refresh("SetRollingCoeff_SPH2SPH"); }
void ChSystemDem::SetRollingCoeff_SPH2WALL(float mu) { m_sys->mu_roll[1] = mu; m_sys->refresh("SetRollingCoeff_SPH2WALL"); }
void ChSystemDem::SetSpinningCoeff_SPH2SPH(float mu) { m_sys->mu_spin[0] = mu; m_sys->refresh("SetSpinningCoeff_SPH2SPH"); }
void ChSystemDem::SetSpinningCoeff_SPH2WALL(float mu) { m_sys->mu_spin[1] = mu; m_sys->refresh("SetSpinningCoeff_SPH2WALL"); }
void ChSystemDem::SetKn_SPH2SPH(double v) { m_sys->Kn[0] = v; m_sys->refresh("SetKn_SPH2SPH"); }
void ChSystemDem::SetKn_SPH2WALL(double v) { m_sys->Kn[1] = v; m_sys->refresh("SetKn_SPH2WALL"); }
void ChSystemDem::SetGn_SPH2SPH(double v) { m_sys->Gn[0] = v; m_sys->refresh("SetGn_SPH2SPH"); }
void ChSystemDem::SetGn_SPH2WALL(double v) { m_sys->Gn[1] = v; m_sys->refresh("SetGn_SPH2WALL"); }
void ChSystemDem::SetKt_SPH2SPH(double v) { m_sys->Kt[0] = v; m_sys->refresh("SetKt_SPH2SPH"); }
void ChSystemDem::SetGt_SPH2SPH(double v) { m_sys->Gt[0] = v; m_sys->refresh("SetGt_SPH2SPH"); }
void ChSystemDem::SetKt_SPH2WALL(double v) { m_sys->Kt[1] = v; m_sys->refresh("SetKt_SPH2WALL"); }
void ChSystemDem::SetGt_SPH2WALL(double v) { m_sys->Gt[1] = v; m_sys->refresh("SetGt_SPH2WALL"); }
void ChSystemDem::SetCohesionRatio(float v) { m_sys->cohesion_over_g = v; m_sys->refresh("SetCohesionRatio"); }
void ChSystemDem::SetAdhesionRatio_SPH2WALL(float v) { m_sys->adhesion_over_g[1] = v; m_sys->refresh("SetAdhesionRatio_SPH2WALL"); }
void ChSystemDem::UseMaterialBasedModel(bool val) {
    const auto old = m_sys->use_mat_based;
    m_sys->use_mat_based = val;
    try {
        m_sys->refresh("UseMaterialBasedModel");
    } catch (...) {
        m_sys->use_mat_based = old;  // the engine refused (not changeable on a running system): keep the mirror consistent with it
        throw;
    }
}
void ChSystemDem::SetYoungModulus_SPH(double v) { m_sys->young[0] = v; m_sys->refresh("SetYoungModulus_SPH"); }
void ChSystemDem::SetYoungModulus_WALL(double v) { m_sys->young[1] = v; m_sys->refresh("SetYoungModulus_WALL"); }
void ChSystemDem::SetPoissonRatio_SPH(double v) { m_sys->poisson[0] = v; m_sys->refresh("SetPoissonRatio_SPH"); }
void ChSystemDem::SetPoissonRatio_WALL(double v) { m_sys->poisson[1] = v; m_sys->refresh("SetPoissonRatio_WALL"); }
void ChSystemDem::SetRestitution_SPH(double v) { m_sys->cor[0] = v; m_sys->refresh("SetRestitution_SPH"); }
void ChSystemDem::SetRestitution_WALL(double v) { m_sys->cor[1] = v; m_sys->refresh("SetRestitution_WALL"); }
void ChSystemDem::SetMaxSafeVelocity_SU(float max_vel) { m_sys->max_safe_vel = max_vel; }
void ChSystemDem::SetPsiFactors(unsigned int psi_T, unsigned int psi_L, float psi_R) {
    m_sys->psi_T = psi_T; m_sys->psi_L = psi_L; m_sys->psi_R = psi_R;  // no simulation-unit system: recorded only
}
void ChSystemDem::SetPsiT(unsigned int psi_T) { m_sys->psi_T = psi_T; }
void ChSystemDem::SetPsiL(unsigned int psi_L) { m_sys->psi_L = psi_L; }
void ChSystemDem::SetPsiR(float psi_R) { m_sys->psi_R = psi_R; }
void ChSystemDem::SetRecordingContactInfo(bool record) {
    m_sys->record_contacts = record;
    if (m_sys->initialized && m_sys->friction == CHDEM_FRICTION_MODE::MULTI_STEP)
        m_sys->check(dem_b200_enable_contact_info(m_sys->h, record ? 1 : 0), "SetRecordingContactInfo");
}
void ChSystemDem::SetSimTime(float time) { m_sys->elapsed = time; }
void ChSystemDem::SetVerbosity(CHDEM_VERBOSITY level) { m_sys->verbosity = level; }
void ChSystemDem::SetParticleDensity(float density) { m_sys->density = density; }
void ChSystemDem::SetParticleRadius(float rad) { m_sys->radius = rad; }

// ---- boundary conditions ----------------------------------------------------------------------------------------------
size_t ChSystemDem::CreateBCPlane(const ChVector3f& pos, const ChVector3f& normal, bool track_forces) {
    if (m_sys->initialized)
        fail("boundary conditions must be created before Initialize");
    BCInfo bc;
    bc.kind = BCKind::PLANE;
    ChVector3d n = ChVector3d(normal).GetNormalized();
    for (int k = 0; k < 3; k++) { bc.pos[k] = pos[k]; bc.normal[k] = n[k]; }
    bc.track_forces = track_forces;
    m_sys->bcs.push_back(bc);
    return m_sys->bcs.size() - 1;
}
size_t ChSystemDem::CreateBCCylinderZ(const ChVector3f& center, float radius, bool outward_normal, bool track_forces) {
    if (m_sys->initialized)
        fail("boundary conditions must be created before Initialize");
    BCInfo bc;
    bc.kind = BCKind::ZCYL;
    for (int k = 0; k < 3; k++) bc.pos[k] = center[k];
    bc.radius = radius;
    bc.spheres_inside = !outward_normal;  // normal pointing outward = obstacle, inward = container
    bc.track_forces = track_forces;
    m_sys->bcs.push_back(bc);
    return m_sys->bcs.size() - 1;
}
// Ball boundary (reference: ChSystemDem_impl.cpp:683-721).  outward_normal == false: obstacle the spheres stay outside of;
// true: cavity.  A ball with mass > 0 is a free body: after every step it is advanced on the host under the reaction force
// of the spheres and gravity (semi-implicit Euler, as ChSystemDem_impl.cpp:867-903; the reaction torque is not used:
// boundary bodies carry no spin in the contact law).
size_t ChSystemDem::CreateBCSphere(const ChVector3f& center, float radius, bool outward_normal, bool track_forces, float mass) {
    if (m_sys->initialized) fail("boundary conditions must be created before Initialize");
    BCInfo bc;
    bc.kind = BCKind::SPHERE;
    bc.pos[0] = center.x(); bc.pos[1] = center.y(); bc.pos[2] = center.z();
    bc.radius = radius;
    bc.spheres_inside = outward_normal;
    bc.mass = mass > 0 ? mass : 0;
    bc.track_forces = track_forces || bc.mass > 0;
    if (bc.mass > 0)
        m_sys->any_offset = true;  // stepwise driver below
    m_sys->bcs.push_back(bc);
    return m_sys->bcs.size() - 1;
}
// Cone about z (reference: ChSystemDem_impl.cpp:723-750): hmax / hmin are world z; outward_normal == false keeps the
// spheres above the surface (a hopper), true below it.
size_t ChSystemDem::CreateBCConeZ(const ChVector3f& tip, float slope, float hmax, float hmin, bool outward_normal, bool track_forces) {
    if (m_sys->initialized) fail("boundary conditions must be created before Initialize");
    BCInfo bc;
    bc.kind = BCKind::ZCONE;
    bc.pos[0] = tip.x(); bc.pos[1] = tip.y(); bc.pos[2] = tip.z();
    bc.slope = slope; bc.hmax = hmax; bc.hmin = hmin;
    bc.spheres_inside = !outward_normal;
    bc.track_forces = track_forces;
    m_sys->bcs.push_back(bc);
    return m_sys->bcs.size() - 1;
}
// The reference registers the plate but no kernel has a force case for BC_type::PLATE: the id is handed out,
// the boundary exerts no force.
size_t ChSystemDem::CreateCustomizedPlate(const ChVector3f& pos_center, const ChVector3f& normal, float /*hdim_y*/) {
    if (m_sys->initialized) fail("boundary conditions must be created before Initialize");
    BCInfo bc;
    bc.kind = BCKind::PLATE;
    bc.pos[0] = pos_center.x(); bc.pos[1] = pos_center.y(); bc.pos[2] = pos_center.z();
    bc.normal[0] = normal.x(); bc.normal[1] = normal.y(); bc.normal[2] = normal.z();
    bc.enabled = false;
    m_sys->bcs.push_back(bc);
    return m_sys->bcs.size() - 1;
}
void ChSystemDem::SetBCSpherePosition(size_t id, const ChVector3f& pos) {
    if (id >= m_sys->bcs.size() || m_sys->bcs[id].kind != BCKind::SPHERE) fail("SetBCSpherePosition: not a sphere boundary");
    BCInfo& bc = m_sys->bcs[id];
    bc.pos[0] = pos.x(); bc.pos[1] = pos.y(); bc.pos[2] = pos.z();
    if (m_sys->initialized)
        m_sys->check(dem_b200_set_wall_state(m_sys->h, bc.wall, bc.pos, bc.vel), "SetBCSpherePosition");
}
void ChSystemDem::SetBCSphereVelocity(size_t id, const ChVector3f& v) {
    if (id >= m_sys->bcs.size() || m_sys->bcs[id].kind != BCKind::SPHERE) fail("SetBCSphereVelocity: not a sphere boundary");
    BCInfo& bc = m_sys->bcs[id];
    bc.vel[0] = v.x(); bc.vel[1] = v.y(); bc.vel[2] = v.z();
    if (m_sys->initialized)
        m_sys->check(dem_b200_set_wall_state(m_sys->h, bc.wall, bc.pos, bc.vel), "SetBCSphereVelocity");
}
ChVector3f ChSystemDem::GetBCSpherePosition(size_t id) const {
    if (id >= m_sys->bcs.size() || m_sys->bcs[id].kind != BCKind::SPHERE) fail("GetBCSpherePosition: not a sphere boundary");
    const BCInfo& bc = m_sys->bcs[id];
    return ChVector3f((float)bc.pos[0], (float)bc.pos[1], (float)bc.pos[2]);
}
ChVector3f ChSystemDem::GetBCSphereVelocity(size_t id) const {
    if (id >= m_sys->bcs.size() || m_sys->bcs[id].kind != BCKind::SPHERE) fail("GetBCSphereVelocity: not a sphere boundary");
    const BCInfo& bc = m_sys->bcs[id];
    return ChVector3f((float)bc.vel[0], (float)bc.vel[1], (float)bc.vel[2]);
}
// Surface spin of a plane boundary (ChSystemDem_impl.cpp:993-1001; used by the force at ChDemBoundaryConditions.cuh:394): the
// plane's material point at a contact moves with omega x (x - center); the plane itself stays where it is.
void ChSystemDem::SetBCPlaneRotation(size_t plane_id, ChVector3d center, ChVector3d omega) {
    ChSystemDem_impl& S = *m_sys;
    if (plane_id >= S.bcs.size() || S.bcs[plane_id].kind != BCKind::PLANE)
        fail("SetBCPlaneRotation: not a plane boundary");
    BCInfo& bc = S.bcs[plane_id];
    for (int k = 0; k < 3; k++) { bc.rot_center[k] = center[k]; bc.rot_omega[k] = omega[k]; }
    bc.has_rotation = true;
    if (S.initialized)
        S.check(dem_b200_set_wall_rotation(S.h, bc.wall, bc.rot_center, bc.rot_omega), "SetBCPlaneRotation");
}
// acceleration of the last step, gravity included (the reference returns its sphere_acc array: ChSystemDem_impl.cpp:1290-1296)
ChVector3f ChSystemDem::GetParticleLinAcc(int i) const {
    const ChSystemDem_impl& S = *m_sys;
    if (!S.initialized)
        return ChVector3f(0, 0, 0);
    double a[3];
    S.check(dem_b200_get_sphere_accel(S.h, (size_t)i, a), "GetParticleLinAcc");
    return ChVector3f((float)a[0], (float)a[1], (float)a[2]);
}
// ---- per-contact records (ChSystemDem_impl.cpp:443-635).  The engine keeps them per contact of the MULTI_STEP contact map;
// like the reference, asking without SetRecordingContactInfo(true) is an error.
namespace {
// info13 of the contact between sphere i and label j (a sphere index, or nSpheres + BC_id + 1: ChDemBoundaryConditions.cuh:102);
// the record is the one of sphere i (for two spheres the partner's is its mirror image)
bool contact_info(const ChSystemDem_impl& S, unsigned int i, unsigned int j, double out[13], const char* what) {
    if (!S.initialized || !S.record_contacts)
        fail(std::string(what) + ": recording_contactInfo set to false (call SetRecordingContactInfo(true))");
    if (S.friction != CHDEM_FRICTION_MODE::MULTI_STEP)
        fail(std::string(what) + ": per-contact records need MULTI_STEP friction");
    const size_t n = S.n();
    if (i >= n && j < n)
        std::swap(i, j);  // sphere-wall: the record sits with the sphere (ChSystemDem_impl.cpp:557-561)
    if (i >= n)
        return false;
    uint32_t other;
    if (j < n) {
        other = (uint32_t)(S.bcs.size() + S.num_triangles() + j);
    } else {
        const size_t bc = (size_t)j - n - 1;
        if (bc >= S.bcs.size() || S.bcs[bc].wall < 0)
            return false;
        other = (uint32_t)S.bcs[bc].wall;
    }
    int found = 0;
    S.check(dem_b200_get_contact_info(S.h, i, other, out, &found), what);
    return found != 0;
}
ChVector3f vec3f(const double* p) { return ChVector3f((float)p[0], (float)p[1], (float)p[2]); }
}  // namespace

ChVector3f ChSystemDem::getNormalForce(unsigned int i, unsigned int j) {
    double c[13];
    return contact_info(*m_sys, i, j, c, "getNormalForce") ? vec3f(c) : ChVector3f(0, 0, 0);
}
ChVector3f ChSystemDem::getSlidingFrictionForce(unsigned int i, unsigned int j) {
    double c[13];
    return contact_info(*m_sys, i, j, c, "getSlidingFrictionForce") ? vec3f(c + 3) : ChVector3f(0, 0, 0);
}
ChVector3f ChSystemDem::getRollingFrictionTorque(unsigned int i, unsigned int j) {
    double c[13];
    if (m_sys->rolling == CHDEM_ROLLING_MODE::NO_RESISTANCE)
        return ChVector3f(0, 0, 0);  // ChSystemDem_impl.cpp:449-451
    return contact_info(*m_sys, i, j, c, "getRollingFrictionTorque") ? vec3f(c + 6) : ChVector3f(0, 0, 0);
}
ChVector3f ChSystemDem::getRollingVrot(unsigned int i, unsigned int j) {
    double c[13];
    if (m_sys->rolling == CHDEM_ROLLING_MODE::NO_RESISTANCE)
        return ChVector3f(0, 0, 0);
    return contact_info(*m_sys, i, j, c, "getRollingVrot") ? vec3f(c + 9) : ChVector3f(0, 0, 0);
}
float ChSystemDem::getRollingCharContactTime(unsigned int i, unsigned int j) {
    double c[13];
    if (m_sys->rolling == CHDEM_ROLLING_MODE::NO_RESISTANCE)
        return 0.f;
    return contact_info(*m_sys, i, j, c, "getRollingCharContactTime") ? (float)c[12] : 0.f;
}

// "bi, bj, n_mag[, fx, fy, fz][, mx, my, mz]", one row per sphere-sphere contact with bi < bj (ChSystemDem_impl.cpp:588-635)
void ChSystemDem::WriteContactInfoFile(const std::string& outfilename) const {
    const ChSystemDem_impl& S = *m_sys;
    if (!S.initialized || !S.record_contacts || S.friction == CHDEM_FRICTION_MODE::FRICTIONLESS)
        fail("WriteContactInfoFile: you did not enable contact info recording or are using the frictionless model");
    if (S.friction != CHDEM_FRICTION_MODE::MULTI_STEP)
        fail("WriteContactInfoFile: per-contact records need MULTI_STEP friction");
    size_t cnt = 0;
    S.check(dem_b200_get_contact_infos(S.h, nullptr, nullptr, nullptr, 0, &cnt), "WriteContactInfoFile");
    std::vector<uint32_t> bi(cnt + 1), bj(cnt + 1);
    std::vector<double> info((cnt + 1) * 13);
    if (cnt)
        S.check(dem_b200_get_contact_infos(S.h, bi.data(), bj.data(), info.data(), cnt, &cnt), "WriteContactInfoFile");
    std::vector<size_t> order(cnt);
    for (size_t k = 0; k < cnt; k++) order[k] = k;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bi[a] != bi[b] ? bi[a] < bi[b] : bj[a] < bj[b]; });
    const bool roll = S.rolling != CHDEM_ROLLING_MODE::NO_RESISTANCE;
    std::ostringstream o;
    o << "bi, bj, n_mag, fx, fy, fz";
    if (roll)
        o << ", mx, my, mz";
    o << "\n";
    for (size_t r = 0; r < cnt; r++) {
        const size_t k = order[r];
        const double* c = &info[13 * k];
        o << bi[k] << ", " << bj[k] << ", " << std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) << ", " << c[3] << ", " << c[4] << ", " << c[5];
        if (roll)
            o << ", " << c[6] << ", " << c[7] << ", " << c[8];
        o << "\n";
    }
    std::ofstream f(outfilename, std::ios::out);
    f << o.str();
}
bool ChSystemDem::DisableBCbyID(size_t id) {
    if (id >= m_sys->bcs.size()) return false;
    m_sys->bcs[id].enabled = false;
    if (m_sys->initialized && m_sys->bcs[id].wall >= 0)
        m_sys->check(dem_b200_enable_wall(m_sys->h, m_sys->bcs[id].wall, 0), "DisableBCbyID");
    return true;
}
bool ChSystemDem::EnableBCbyID(size_t id) {
    if (id >= m_sys->bcs.size()) return false;
    m_sys->bcs[id].enabled = true;
    if (m_sys->initialized && m_sys->bcs[id].wall >= 0)
        m_sys->check(dem_b200_enable_wall(m_sys->h, m_sys->bcs[id].wall, 1), "EnableBCbyID");
    return true;
}
bool ChSystemDem::SetBCOffsetFunction(size_t id, const GranPositionFunction& f) {
    if (id >= m_sys->bcs.size()) return false;
    m_sys->bcs[id].offset = f;
    m_sys->bcs[id].has_offset = true;
    m_sys->any_offset = true;
    return true;
}
void ChSystemDem::setBDWallsMotionFunction(const GranPositionFunction& pos_fn) {
    for (size_t id = 0; id < NUM_RESERVED_BC_IDS; id++)
        SetBCOffsetFunction(id, pos_fn);
}

// ---- run --------------------------------------------------------------------------------------------------------------
void ChSystemDem::Initialize() {
    ChSystemDem_impl& S = *m_sys;
    if (S.initialized)
        fail("Initialize called twice");
    const size_t n = S.pos.size() / 3;
    if (n == 0)
        fail("Initialize without particles");
    if (S.rad.size() != n)
        S.rad.assign(n, S.radius);
    if (S.fixed.size() != n)
        S.fixed.resize(n, 0);
    dem_b200_config cfg = S.make_config();
    int rc = dem_b200_create(&cfg, &S.h);
    if (rc < 0)
        fail(std::string("no usable CUDA device: ") + dem_b200_last_error(nullptr));
    // six reserved big-domain planes, normals pointing inwards (ChSystemDem_impl.cpp:102-132)
    const double c[3] = {S.O[0], S.O[1], S.O[2]};
    for (int a = 0; a < 3; a++)
        for (int side = 0; side < 2; side++) {
            BCInfo& bc = S.bcs[2 * a + side];
            bc.kind = BCKind::PLANE;
            for (int k = 0; k < 3; k++) { bc.pos[k] = 0; bc.normal[k] = 0; }
            bc.pos[a] = c[a] + (side == 0 ? -0.5 : 0.5) * S.box[a];
            bc.normal[a] = side == 0 ? 1.0 : -1.0;
        }
    bool track = false;
    for (auto& bc : S.bcs) {
        if (bc.kind == BCKind::PLANE || bc.kind == BCKind::PLATE)
            bc.wall = dem_b200_add_plane_wall(S.h, bc.pos, bc.normal);  // a PLATE is created disabled: no force case upstream
        else if (bc.kind == BCKind::SPHERE)
            bc.wall = dem_b200_add_sphere_wall(S.h, bc.pos, bc.radius, bc.spheres_inside ? 0 : 1);
        else if (bc.kind == BCKind::ZCONE)
            bc.wall = dem_b200_add_zcone_wall(S.h, bc.pos, bc.slope, bc.hmin, bc.hmax, bc.spheres_inside ? 1 : 0);
        else
            bc.wall = dem_b200_add_zcylinder_wall(S.h, bc.pos, bc.radius, bc.spheres_inside ? 1 : 0);
        if (bc.wall < 0)
            fail(std::string("CreateBC: ") + dem_b200_last_error(S.h) +
                 " (the engine holds at most 16 boundary conditions, the six reserved box planes included)");
        if (bc.has_rotation)
            S.check(dem_b200_set_wall_rotation(S.h, bc.wall, bc.rot_center, bc.rot_omega), "SetBCPlaneRotation");
        if (!bc.enabled)
            S.check(dem_b200_enable_wall(S.h, bc.wall, 0), "DisableBCbyID");
        track |= bc.track_forces;
    }
    if (track)
        S.check(dem_b200_track_wall_forces(S.h, 1), "track_forces");
    for (int k = 0; k < 3; k++) {
        dem_b200_contact_class cc = S.make_class(k);
        S.check(dem_b200_set_contact_class(S.h, k, &cc), "set_contact_class");
    }
    for (size_t m = 0; m < S.meshes.size(); m++) {
        auto& M = S.meshes[m];
        const int id = dem_b200_add_mesh(S.h, M.verts9.size() / 9, M.verts9.data(), M.mass);
        S.check(id, "AddMesh");
        if (M.has_motion)
            S.check(dem_b200_set_mesh_motion(S.h, id, M.pos, M.rot, M.lin, M.ang), "ApplyMeshMotion");
    }
    if (!S.meshes.empty() && !S.mesh_collision)
        S.check(dem_b200_enable_mesh_collision(S.h, 0), "EnableMeshCollision");
    S.check(dem_b200_set_spheres(S.h, n, S.pos.data(), S.vel.data(), S.omg.data(), S.rad.data(), S.fixed.data()),
            "SetParticles");
    {
        // engine shape ids: wall w is shape w (walls are added in BC-id order), then the mesh triangles, then sphere i
        // is shape nW + nT + i
        const uint32_t nW = (uint32_t)S.bcs.size();
        const uint32_t base = nW + S.num_triangles();
        for (auto& r : S.hist) {
            if (r.is_bc && r.partner >= nW) {
                // label nSpheres + 1 + nBCs + 1 + family: a mesh-family record (ChDemSMCtrimesh.cu:728) cannot be
                // attached to one facet; the contact restarts with an empty history
                if (!S.meshes.empty() && r.partner > nW)
                    continue;
                fail("contact history refers to a boundary condition that was not created before Initialize");
            }
            S.check(dem_b200_add_history(S.h, base + r.sphere, r.is_bc ? r.partner : base + r.partner, r.d, 0.0, 0.0),
                    "ReadContactHistory");
        }
    }
    S.check(dem_b200_initialize(S.h), "Initialize");
    S.initialized = true;
    if (S.record_contacts && S.friction == CHDEM_FRICTION_MODE::MULTI_STEP)
        S.check(dem_b200_enable_contact_info(S.h, 1), "SetRecordingContactInfo");
    if (S.verbosity != CHDEM_VERBOSITY::QUIET)
        printf("ChSystemDem (B200 engine): %zu spheres, %zu boundary conditions, h = %g\n", n, S.bcs.size(), (double)S.step);
}

double ChSystemDem::AdvanceSimulation(float duration) {
    ChSystemDem_impl& S = *m_sys;
    if (!S.initialized)
        fail("AdvanceSimulation before Initialize");
    const auto t0 = std::chrono::steady_clock::now();
    const int nsteps = (int)std::lround((double)duration / S.step);  // ChDemSMC.cu:623-624
    if (!S.any_offset) {
        S.check(dem_b200_step(S.h, nsteps), "AdvanceSimulation");
    } else {
        // moving boundaries: position at the start of each step, velocity = midpoint difference quotient
        // (ChSystemDem_impl.cpp:863-948)
        for (int i = 0; i < nsteps; i++) {
            const float t = (float)(S.elapsed + (double)i * S.step);
            for (auto& bc : S.bcs) {
                if (!bc.has_offset || (S.BD_fixed && (&bc - &S.bcs[0]) < (long)NUM_RESERVED_BC_IDS))
                    continue;
                const double3 o0 = bc.offset(t), o1 = bc.offset(t + S.step);
                double p[3] = {bc.pos[0] + o0.x, bc.pos[1] + o0.y, bc.pos[2] + o0.z};
                double v[3] = {(o1.x - o0.x) / S.step, (o1.y - o0.y) / S.step, (o1.z - o0.z) / S.step};
                S.check(dem_b200_set_wall_state(S.h, bc.wall, p, v), "SetBCOffsetFunction");
            }
            S.check(dem_b200_step(S.h, 1), "AdvanceSimulation");
            for (auto& bc : S.bcs) {
                if (bc.kind != BCKind::SPHERE || !(bc.mass > 0) || !bc.enabled)
                    continue;
                double f[3];
                S.check(dem_b200_wall_force(S.h, bc.wall, f), "BC sphere reaction force");  // synchronises: one step at a time
                for (int k = 0; k < 3; k++) {
                    bc.vel[k] += (f[k] / bc.mass + S.grav[k]) * S.step;
                    bc.pos[k] += bc.vel[k] * S.step;
                }
                S.check(dem_b200_set_wall_state(S.h, bc.wall, bc.pos, bc.vel), "BC sphere motion");
            }
        }
    }
    S.elapsed += (double)nsteps * S.step;
    if (S.verbosity == CHDEM_VERBOSITY::METRICS) {
        S.check(dem_b200_sync(S.h), "AdvanceSimulation");
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        m_RTF = (float)(wall / std::max(1e-30, (double)nsteps * S.step));
    }
    return (double)nsteps * S.step;
}

// ---- state access -----------------------------------------------------------------------------------------------------
void ChSystemDem::SetParticlePosition(int i, const ChVector3d p) {
    ChSystemDem_impl& S = *m_sys;
    if (!S.initialized) {
        for (int k = 0; k < 3; k++) S.pos[3 * (size_t)i + k] = p[k];
        return;
    }
    const size_t n = S.n();
    std::vector<double> all(3 * n);
    S.check(dem_b200_get_state(S.h, all.data(), nullptr, nullptr), "SetParticlePosition");
    for (int k = 0; k < 3; k++) all[3 * (size_t)i + k] = p[k];
    S.check(dem_b200_set_state(S.h, all.data(), nullptr, nullptr), "SetParticlePosition");
}
void ChSystemDem::SetParticleVelocity(int i, const ChVector3d v) {
    ChSystemDem_impl& S = *m_sys;
    if (!S.initialized) {
        for (int k = 0; k < 3; k++) S.vel[3 * (size_t)i + k] = v[k];
        return;
    }
    const size_t n = S.n();
    std::vector<double> all(3 * n);
    S.check(dem_b200_get_state(S.h, nullptr, all.data(), nullptr), "SetParticleVelocity");
    for (int k = 0; k < 3; k++) all[3 * (size_t)i + k] = v[k];
    S.check(dem_b200_set_state(S.h, nullptr, all.data(), nullptr), "SetParticleVelocity");
}

float ChSystemDem::GetSimTime() const { return (float)m_sys->elapsed; }
size_t ChSystemDem::GetNumParticles() const { return m_sys->n(); }
float ChSystemDem::GetParticleRadius() const { return m_sys->radius; }
bool ChSystemDem::IsFixed(int i) const { return m_sys->fixed.at((size_t)i) != 0; }
unsigned int ChSystemDem::GetNumSDs() const { return 0; }  // no subdomain partition in this engine
size_t ChSystemDem::EstimateMemUsage() const { return m_sys->n() * 1900; }

static double reduce(const ChSystemDem_impl& S, int which, double arg, const char* what) {
    if (!S.initialized)
        fail(std::string(what) + " before Initialize");
    double out = 0;
    S.check(dem_b200_reduce(S.h, which, arg, &out), what);
    return out;
}
double ChSystemDem::GetMaxParticleZ() const { return reduce(*m_sys, DEMB200_RED_MAX_Z, 0, "GetMaxParticleZ"); }
double ChSystemDem::GetMinParticleZ() const { return reduce(*m_sys, DEMB200_RED_MIN_Z, 0, "GetMinParticleZ"); }
unsigned int ChSystemDem::GetNumParticleAboveZ(float z) const {
    return (unsigned int)reduce(*m_sys, DEMB200_RED_COUNT_ABOVE_Z, z, "GetNumParticleAboveZ");
}
unsigned int ChSystemDem::GetNumParticleAboveX(float x) const {
    return (unsigned int)reduce(*m_sys, DEMB200_RED_COUNT_ABOVE_X, x, "GetNumParticleAboveX");
}
float ChSystemDem::GetParticlesKineticEnergy() const {
    // the reference sums m v^2 / 2 only (ChSystemDem_impl.cpp:1250-1264); the engine adds the rotational part
    return (float)reduce(*m_sys, DEMB200_RED_KE_TRANSLATIONAL, 0, "GetParticlesKineticEnergy");
}
unsigned int ChSystemDem::GetNumContacts() const {
    if (m_sys->friction != CHDEM_FRICTION_MODE::MULTI_STEP)
        return 0;
    return (unsigned int)(reduce(*m_sys, DEMB200_RED_NUM_CONTACTS, 0, "GetNumContacts") / 2);
}

static ChVector3f get3(const ChSystemDem_impl& S, int i, int what) {
    double p[3], v[3], w[3];
    if (!S.initialized) {
        const std::vector<double>& a = what == 0 ? S.pos : what == 1 ? S.vel : S.omg;
        return ChVector3f((float)a[3 * (size_t)i], (float)a[3 * (size_t)i + 1], (float)a[3 * (size_t)i + 2]);
    }
    S.check(dem_b200_get_sphere(S.h, (size_t)i, p, v, w), "GetParticle*");
    const double* a = what == 0 ? p : what == 1 ? v : w;
    return ChVector3f((float)a[0], (float)a[1], (float)a[2]);
}
ChVector3f ChSystemDem::GetParticlePosition(int i) const { return get3(*m_sys, i, 0); }
ChVector3f ChSystemDem::GetParticleVelocity(int i) const { return get3(*m_sys, i, 1); }
ChVector3f ChSystemDem::GetParticleAngVelocity(int i) const {
    if (m_sys->friction == CHDEM_FRICTION_MODE::FRICTIONLESS)
        return ChVector3f(0);
    return get3(*m_sys, i, 2);
}
ChVector3f ChSystemDem::GetBCPlanePosition(size_t id) const {
    const BCInfo& bc = m_sys->bcs.at(id);
    double3 o = bc.has_offset ? bc.offset((float)m_sys->elapsed) : make_double3(0, 0, 0);
    return ChVector3f((float)(bc.pos[0] + o.x), (float)(bc.pos[1] + o.y), (float)(bc.pos[2] + o.z));
}
bool ChSystemDem::GetBCReactionForces(size_t id, ChVector3f& force) const {
    if (id >= m_sys->bcs.size() || !m_sys->bcs[id].track_forces || !m_sys->initialized)
        return false;
    double f[3];
    m_sys->check(dem_b200_wall_force(m_sys->h, m_sys->bcs[id].wall, f), "GetBCReactionForces");
    force = ChVector3f((float)f[0], (float)f[1], (float)f[2]);
    return true;
}

// ---- output -----------------------------------------------------------------------------------------------------------
namespace {
struct Snapshot {
    std::vector<double> pos, vel, omg;
};
Snapshot snapshot(const ChSystemDem_impl& S) {
    Snapshot s;
    const size_t n = S.n();
    if (!S.initialized) {
        s.pos = S.pos; s.vel = S.vel; s.omg = S.omg;
        return s;
    }
    s.pos.resize(3 * n); s.vel.resize(3 * n); s.omg.resize(3 * n);
    S.check(dem_b200_get_state(S.h, s.pos.data(), s.vel.data(), s.omg.data()), "WriteParticleFile");
    return s;
}
}  // namespace

// CSV layout of the reference: ChSystemDem_impl.cpp:259-333 (header, then one row per sphere, default ostream floats)
void ChSystemDem::WriteCsvParticles(std::ofstream& ptFile) const {
    const ChSystemDem_impl& S = *m_sys;
    const Snapshot s = snapshot(S);
    const unsigned f = S.out_flags;
    const bool fr = S.friction != CHDEM_FRICTION_MODE::FRICTIONLESS;
    std::ostringstream o;
    o << "x,y,z";
    if (f & VEL_COMPONENTS) o << ",vx,vy,vz";
    if (f & ABSV) o << ",absv";
    if (f & FIXITY) o << ",fixed";
    if (fr && (f & ANG_VEL_COMPONENTS)) o << ",wx,wy,wz";
    if (f & FORCE_COMPONENTS) o << ",fx,fy,fz";
    o << "\n";
    // contact force on each sphere = (acceleration of the last step - g) * mass (ChSystemDem_impl.cpp:322-327)
    std::vector<double> acc;
    if (f & FORCE_COMPONENTS) {
        acc.assign(3 * S.n(), 0.0);
        if (S.initialized)
            S.check(dem_b200_get_accel(S.h, acc.data()), "WriteParticleFile");
    }
    for (size_t i = 0; i < S.n(); i++) {
        const float x = (float)s.pos[3 * i], y = (float)s.pos[3 * i + 1], z = (float)s.pos[3 * i + 2];
        o << x << "," << y << "," << z;
        const float vx = (float)s.vel[3 * i], vy = (float)s.vel[3 * i + 1], vz = (float)s.vel[3 * i + 2];
        if (f & VEL_COMPONENTS) o << "," << vx << "," << vy << "," << vz;
        if (f & ABSV) o << "," << (float)std::sqrt((double)vx * vx + (double)vy * vy + (double)vz * vz);
        if (f & FIXITY) o << "," << (int)S.fixed[i];
        if (fr && (f & ANG_VEL_COMPONENTS))
            o << "," << (float)s.omg[3 * i] << "," << (float)s.omg[3 * i + 1] << "," << (float)s.omg[3 * i + 2];
        if (f & FORCE_COMPONENTS) {
            const double r = S.rad.size() == S.n() ? S.rad[i] : (double)S.radius;
            const double m = 4.0 / 3.0 * 3.14159265358979323846 * r * r * r * S.density;
            for (int k = 0; k < 3; k++)  // before the first step the acceleration is zero, as the reference's sphere_acc
                o << "," << (acc[3 * i + k] - (double)S.grav[k]) * m;
        }
        o << "\n";
    }
    ptFile << o.str();
}

// headerless packed fp32 records: ChSystemDem_impl.cpp:212-257
void ChSystemDem::WriteRawParticles(std::ofstream& ptFile) const {
    const ChSystemDem_impl& S = *m_sys;
    const Snapshot s = snapshot(S);
    const unsigned f = S.out_flags;
    const bool fr = S.friction != CHDEM_FRICTION_MODE::FRICTIONLESS;
    std::vector<float> rec;
    for (size_t i = 0; i < S.n(); i++) {
        rec.clear();
        for (int k = 0; k < 3; k++) rec.push_back((float)s.pos[3 * i + k]);
        const float vx = (float)s.vel[3 * i], vy = (float)s.vel[3 * i + 1], vz = (float)s.vel[3 * i + 2];
        if (f & VEL_COMPONENTS) { rec.push_back(vx); rec.push_back(vy); rec.push_back(vz); }
        if (f & ABSV) rec.push_back((float)std::sqrt((double)vx * vx + (double)vy * vy + (double)vz * vz));
        if (fr && (f & ANG_VEL_COMPONENTS))
            for (int k = 0; k < 3; k++) rec.push_back((float)s.omg[3 * i + k]);
        ptFile.write((const char*)rec.data(), (std::streamsize)(rec.size() * sizeof(float)));
    }
}

void ChSystemDem::WriteParticleFile(const std::string& outfilename) const {
    switch (m_sys->out_mode) {
        case CHDEM_OUTPUT_MODE::NONE:
            return;
        case CHDEM_OUTPUT_MODE::BINARY: {
            std::ofstream f(outfilename, std::ios::out | std::ios::binary);
            WriteRawParticles(f);
            return;
        }
        case CHDEM_OUTPUT_MODE::HDF5:
            fail("HDF5 output is not available in this build (the reference needs USE_HDF5 too, ChSystemDem_impl.cpp:336)");
        default: {
            std::ofstream f(outfilename, std::ios::out);
            WriteCsvParticles(f);
        }
    }
}

// ---- checkpoint (text grammar of the reference: ChSystemDem.cpp:1322-1412, 1450-1500; reader :704-857, 900-1169) ----
void ChSystemDem::WriteCheckpointParams(std::ofstream& cp) const {
    const ChSystemDem_impl& S = *m_sys;
    std::ostringstream p;
    p << "nSpheres: " << GetNumParticles() << "\n";
    p << "density: " << S.density << "\n";
    p << "radius: " << S.radius << "\n";
    p << "boxSize: " << S.box[0] << " " << S.box[1] << " " << S.box[2] << "\n";
    p << "BDFixed: " << (int)S.BD_fixed << "\n";
    p << "BDCenter: " << S.O[0] << " " << S.O[1] << " " << S.O[2] << "\n";
    p << "verbosity: " << (unsigned)S.verbosity << "\n";
    p << "useMinLengthUnit: " << (int)S.use_min_length << "\n";
    p << "recordContactInfo: " << (int)S.record_contacts << "\n";
    p << "particleFileMode: " << (unsigned)S.out_mode << "\n";
    p << "particleFileFlags: " << S.out_flags << "\n";
    p << "fixedStepSize: " << S.step << "\n";
    p << "cohesionOverG: " << (float)S.cohesion_over_g << "\n";
    p << "adhesionOverG_s2w: " << (float)S.adhesion_over_g[1] << "\n";
    p << "G: " << S.grav[0] << " " << S.grav[1] << " " << S.grav[2] << "\n";
    p << "elapsedTime: " << GetSimTime() << "\n";
    p << "K_n_s2s: " << S.Kn[0] << "\n" << "K_n_s2w: " << S.Kn[1] << "\n";
    p << "K_t_s2s: " << S.Kt[0] << "\n" << "K_t_s2w: " << S.Kt[1] << "\n";
    p << "G_n_s2s: " << S.Gn[0] << "\n" << "G_n_s2w: " << S.Gn[1] << "\n";
    p << "G_t_s2s: " << S.Gt[0] << "\n" << "G_t_s2w: " << S.Gt[1] << "\n";
    p << "RollingCoeff_s2s: " << S.mu_roll[0] << "\n" << "RollingCoeff_s2w: " << S.mu_roll[1] << "\n";
    p << "SpinningCoeff_s2s: " << S.mu_spin[0] << "\n" << "SpinningCoeff_s2w: " << S.mu_spin[1] << "\n";
    p << "StaticFrictionCoeff_s2s: " << S.mu[0] << "\n" << "StaticFrictionCoeff_s2w: " << S.mu[1] << "\n";
    p << "PsiT: " << S.psi_T << "\n" << "PsiL: " << S.psi_L << "\n" << "PsiR: " << S.psi_R << "\n";
    p << "frictionMode: " << (unsigned)S.friction << "\n";
    p << "rollingMode: " << (unsigned)S.rolling << "\n";
    p << "timeIntegrator: " << (unsigned)S.integrator << "\n";
    p << "maxSafeVelSU: " << S.max_safe_vel << "\n";
    cp << p.str();
}

// "partners 12 history 12", then per sphere 12 partner labels and 12 x 3 history floats.  Labels: sphere index, or
// nSpheres + BC_id + 1 for a boundary (ChDemBoundaryConditions.cuh:102), NULL_CHDEM_ID for an empty slot.  Each sphere
// lists every contact it takes part in (as the reference does); the displacement is written from that sphere's point of
// view (sign flipped for the higher-id partner, see the reader).
using PartnerRows = std::vector<std::vector<std::pair<uint32_t, std::array<float, 3>>>>;
static PartnerRows partner_rows(const ChSystemDem_impl& S) {
    const size_t n = S.n();
    PartnerRows rows(n);
    if (S.initialized && S.friction == CHDEM_FRICTION_MODE::MULTI_STEP) {
        size_t m = 0;
        S.check(dem_b200_get_history(S.h, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &m), "WriteHstHistory");
        std::vector<uint32_t> owner(m), other(m);
        std::vector<double> d(3 * m);
        if (m)
            S.check(dem_b200_get_history(S.h, owner.data(), other.data(), d.data(), nullptr, nullptr, m, &m), "WriteHstHistory");
        const uint32_t nW = (uint32_t)dem_b200_num_walls(S.h);
        const uint32_t base = nW + (uint32_t)dem_b200_num_triangles(S.h);
        for (size_t c = 0; c < m; c++) {
            const uint32_t so = owner[c] - base;  // owner = higher shape id = body 2 of the canonical orientation
            std::array<float, 3> hv = {(float)-d[3 * c], (float)-d[3 * c + 1], (float)-d[3 * c + 2]};
            if (other[c] < nW) {  // wall: label by BC id
                size_t bc = 0;
                for (; bc < S.bcs.size(); bc++)
                    if (S.bcs[bc].wall == (int)other[c]) break;
                rows[so].push_back({(uint32_t)(n + bc + 1), hv});
            } else if (other[c] < base) {
                // mesh facet: the reference keeps one record per (sphere, mesh family), labelled
                // nSpheres + 1 + nBCs + 1 + family (ChDemSMCtrimesh.cu:728); write the first facet's record of a family
                uint32_t tri = other[c] - nW, fam = 0;
                for (; fam < S.meshes.size() && tri >= S.meshes[fam].verts9.size() / 9; fam++)
                    tri -= (uint32_t)(S.meshes[fam].verts9.size() / 9);
                const uint32_t label = (uint32_t)(n + 1 + S.bcs.size() + 1 + fam);
                bool seen = false;
                for (auto& r : rows[so]) seen |= (r.first == label);
                if (!seen)
                    rows[so].push_back({label, hv});
            } else {
                const uint32_t sp = other[c] - base;
                rows[so].push_back({sp, hv});
                rows[sp].push_back({so, {-hv[0], -hv[1], -hv[2]}});
            }
        }
    }
    return rows;
}

void ChSystemDem::WriteHstHistory(std::ofstream& hf) const {
    const ChSystemDem_impl& S = *m_sys;
    const size_t n = S.n();
    const unsigned K = MAX_SPHERES_TOUCHED_BY_SPHERE;
    // Layout of the reference writer (ChSystemDem.cpp:1450-1500): FRICTIONLESS writes nothing, SINGLE_STEP the partner map
    // only ("partners 12"), MULTI_STEP the partner map and the displacement triples ("partners 12 history 12").
    if (S.friction == CHDEM_FRICTION_MODE::FRICTIONLESS) {
        printf("WARNING! Currently using FRICTIONLESS model. There is no contact history to write!\n");
        return;
    }
    const bool with_hist = S.friction == CHDEM_FRICTION_MODE::MULTI_STEP;
    const PartnerRows rows = partner_rows(S);
    std::ostringstream o;
    o << "partners " << K;
    if (with_hist)
        o << " history " << K;
    o << "\n";
    for (size_t i = 0; i < n; i++) {
        for (unsigned k = 0; k < K; k++)
            o << (k < rows[i].size() ? rows[i][k].first : (uint32_t)NULL_CHDEM_ID) << " ";
        for (unsigned k = 0; with_hist && k < K; k++) {
            if (k < rows[i].size())
                o << rows[i][k].second[0] << " " << rows[i][k].second[1] << " " << rows[i][k].second[2] << " ";
            else
                o << "0 0 0 ";
        }
        o << "\n";
    }
    hf << o.str();
}

// partner labels of one sphere, as the reference's contact_partners_map row (ChSystemDem_impl.cpp:469-480): sphere
// indices, nSpheres + BC_id + 1 for boundaries, the mesh-family label for facets.  MULTI_STEP friction only (the engine
// keeps a contact map only where there is history to keep).
void ChSystemDem::getNeighbors(unsigned int ID, std::vector<unsigned int>& neighborList) {
    const ChSystemDem_impl& S = *m_sys;
    if (ID >= S.n())
        fail("getNeighbors: bad sphere id");
    const PartnerRows rows = partner_rows(S);
    for (auto& r : rows[ID])
        neighborList.push_back(r.first);
}

void ChSystemDem::WriteContactHistoryFile(const std::string& outfilename) const {
    std::ofstream f(outfilename, std::ios::out);
    WriteHstHistory(f);
}

void ChSystemDem::WriteCheckpointFile(const std::string& outfilename) {
    std::ofstream cp(outfilename, std::ios::out);
    cp << "ChSystemDem\n";
    WriteCheckpointParams(cp);
    cp << "ParamsEnd\n\n";
    cp << "CsvParticles\n";
    const unsigned int flags = m_sys->out_flags;
    const CHDEM_FRICTION_MODE fm = m_sys->friction;
    m_sys->out_flags = VEL_COMPONENTS | FIXITY | ANG_VEL_COMPONENTS;
    WriteCsvParticles(cp);
    m_sys->out_flags = flags;
    cp << "\n";
    if (fm != CHDEM_FRICTION_MODE::FRICTIONLESS) {
        cp << "HstHistory\n";
        WriteHstHistory(cp);
        cp << "\n";
    }
}

bool ChSystemDem::SetParamsFromIdentifier(const std::string& id, std::istringstream& iss, bool /*overwrite*/) {
    ChSystemDem_impl& S = *m_sys;
    unsigned u;
    if (id == "density") iss >> S.density;
    else if (id == "radius") iss >> S.radius;
    else if (id == "boxSize") iss >> S.box[0] >> S.box[1] >> S.box[2];
    else if (id == "BDFixed") { iss >> u; S.BD_fixed = u != 0; }
    else if (id == "BDCenter") iss >> S.O[0] >> S.O[1] >> S.O[2];
    else if (id == "verbosity") { iss >> u; S.verbosity = (CHDEM_VERBOSITY)u; }
    else if (id == "useMinLengthUnit") { iss >> u; S.use_min_length = u != 0; }
    else if (id == "recordContactInfo") { iss >> u; S.record_contacts = u != 0; }
    else if (id == "particleFileMode") { iss >> u; S.out_mode = (CHDEM_OUTPUT_MODE)u; }
    else if (id == "particleFileFlags") iss >> S.out_flags;
    else if (id == "fixedStepSize") iss >> S.step;
    else if (id == "cohesionOverG") iss >> S.cohesion_over_g;
    else if (id == "adhesionOverG_s2w") iss >> S.adhesion_over_g[1];
    else if (id == "G") iss >> S.grav[0] >> S.grav[1] >> S.grav[2];
    else if (id == "elapsedTime") iss >> S.elapsed;
    else if (id == "K_n_s2s") iss >> S.Kn[0];
    else if (id == "K_n_s2w") iss >> S.Kn[1];
    else if (id == "K_t_s2s") iss >> S.Kt[0];
    else if (id == "K_t_s2w") iss >> S.Kt[1];
    else if (id == "G_n_s2s") iss >> S.Gn[0];
    else if (id == "G_n_s2w") iss >> S.Gn[1];
    else if (id == "G_t_s2s") iss >> S.Gt[0];
    else if (id == "G_t_s2w") iss >> S.Gt[1];
    else if (id == "RollingCoeff_s2s") iss >> S.mu_roll[0];
    else if (id == "RollingCoeff_s2w") iss >> S.mu_roll[1];
    else if (id == "SpinningCoeff_s2s") iss >> S.mu_spin[0];
    else if (id == "SpinningCoeff_s2w") iss >> S.mu_spin[1];
    else if (id == "StaticFrictionCoeff_s2s") iss >> S.mu[0];
    else if (id == "StaticFrictionCoeff_s2w") iss >> S.mu[1];
    else if (id == "PsiT") iss >> S.psi_T;
    else if (id == "PsiL") iss >> S.psi_L;
    else if (id == "PsiR") iss >> S.psi_R;
    else if (id == "frictionMode") { iss >> u; S.friction = (CHDEM_FRICTION_MODE)u; }
    else if (id == "rollingMode") { iss >> u; S.rolling = (CHDEM_ROLLING_MODE)u; }
    else if (id == "timeIntegrator") { iss >> u; S.integrator = (CHDEM_TIME_INTEGRATOR)u; }
    else if (id == "maxSafeVelSU") iss >> S.max_safe_vel;
    else if (id == "nSpheres") { /* consumed by ReadDatParams */ }
    else return false;
    return true;
}

unsigned int ChSystemDem::ReadDatParams(std::ifstream& ifile, bool overwrite) {
    std::string line;
    unsigned int nSpheres = 0;
    while (std::getline(ifile, line)) {
        if (line.find("ParamsEnd") != std::string::npos)
            break;
        const size_t colon = line.find(':');
        if (colon == std::string::npos)
            continue;
        const std::string id = line.substr(0, colon);
        std::istringstream iss(line.substr(colon + 1));
        if (id == "nSpheres") {
            std::istringstream t(line.substr(colon + 1));
            t >> nSpheres;
        }
        if (!SetParamsFromIdentifier(id, iss, overwrite) && m_sys->verbosity != CHDEM_VERBOSITY::QUIET)
            printf("ChSystemDem: unknown checkpoint parameter \"%s\" skipped\n", id.c_str());
    }
    return nSpheres;
}

void ChSystemDem::ReadCsvParticles(std::ifstream& ifile, unsigned int totRow) {
    ChSystemDem_impl& S = *m_sys;
    std::string line;
    if (!std::getline(ifile, line))
        fail("particle file: missing header");
    std::vector<std::string> cols;
    {
        std::istringstream h(line);
        std::string c;
        while (std::getline(h, c, ',')) {
            while (!c.empty() && (c.back() == '\r' || c.back() == ' ')) c.pop_back();
            cols.push_back(c);
        }
    }
    auto col = [&](const char* name) {
        for (size_t i = 0; i < cols.size(); i++)
            if (cols[i] == name) return (int)i;
        return -1;
    };
    const int ix = col("x"), iy = col("y"), iz = col("z"), ivx = col("vx"), ivy = col("vy"), ivz = col("vz"),
              ifx = col("fixed"), iwx = col("wx"), iwy = col("wy"), iwz = col("wz");
    if (ix < 0 || iy < 0 || iz < 0)
        fail("particle file: x,y,z columns are required");
    S.pos.clear(); S.vel.clear(); S.omg.clear(); S.fixed.clear();
    unsigned int rows = 0;
    while (rows < totRow && std::getline(ifile, line)) {
        if (line.empty() || line == "\r")
            break;
        std::vector<double> v;
        std::istringstream r(line);
        std::string c;
        while (std::getline(r, c, ','))
            v.push_back(std::stod(c));
        auto get = [&](int i) { return (i >= 0 && i < (int)v.size()) ? v[i] : 0.0; };
        S.pos.push_back(get(ix)); S.pos.push_back(get(iy)); S.pos.push_back(get(iz));
        S.vel.push_back(get(ivx)); S.vel.push_back(get(ivy)); S.vel.push_back(get(ivz));
        S.omg.push_back(get(iwx)); S.omg.push_back(get(iwy)); S.omg.push_back(get(iwz));
        S.fixed.push_back(get(ifx) != 0 ? 1 : 0);
        rows++;
    }
    S.rad.assign(rows, S.radius);
}

void ChSystemDem::ReadHstHistory(std::ifstream& ifile, unsigned int totItem) {
    ChSystemDem_impl& S = *m_sys;
    std::string line, w1, w2;
    auto blank = [](const std::string& l) { return l.find_first_not_of(" \r\t") == std::string::npos; };
    // header as the reference parses it (quarryHistoryFormat, ChSystemDem.cpp:690-706): "partners N", optionally followed by
    // "history N" (SINGLE_STEP checkpoints carry the partner map only); blank lines in front of it are skipped (:793-795)
    do {
        if (!std::getline(ifile, line))
            fail("history: missing header");
    } while (blank(line));
    unsigned np = 0, nh = 0;
    bool with_hist = false;
    {
        std::istringstream h(line);
        h >> w1 >> np;
        if (w1 != "partners" || np == 0)
            fail("history: header must read \"partners N [history N]\"");
        if (h >> w2 >> nh) {
            if (w2 != "history" || nh != np)
                fail("history: header must read \"partners N [history N]\"");
            with_hist = true;
        }
    }
    const size_t n = S.n();
    S.hist.clear();
    for (size_t i = 0; i < n && i < totItem;) {
        if (!std::getline(ifile, line))
            break;
        if (blank(line))
            continue;  // as the reference (:815-817)
        i++;
        if (!with_hist)
            continue;  // partner map only: nothing to restore, contact pairs are found again by the first step
        const size_t row_i = i - 1;
        std::istringstream r(line);
        std::vector<uint32_t> partner(np);
        for (auto& p : partner) r >> p;
        for (unsigned k = 0; k < np; k++) {
            double d[3];
            r >> d[0] >> d[1] >> d[2];
            const uint32_t p = partner[k];
            if (p == (uint32_t)NULL_CHDEM_ID)
                continue;
            ChSystemDem_impl::HistRow row;
            row.sphere = (uint32_t)row_i;
            if (p > n) {  // boundary label nSpheres + BC_id + 1 (ChDemBoundaryConditions.cuh:102)
                row.is_bc = true;
                row.partner = p - (uint32_t)n - 1;
            } else {
                if (p > row_i)
                    continue;  // the higher-id partner's copy is the one we keep (both are present in the file)
                row.is_bc = false;
                row.partner = p;
            }
            // canonical orientation: body 1 = wall / lower id; the file holds the view of sphere i = body 2
            for (int c = 0; c < 3; c++) row.d[c] = -d[c];
            S.hist.push_back(row);
        }
    }
}

void ChSystemDem::ReadParticleFile(const std::string& infilename) {
    std::ifstream f(infilename);
    if (!f)
        fail("cannot open " + infilename);
    ReadCsvParticles(f);
    m_sys->defragment = false;
}
void ChSystemDem::ReadContactHistoryFile(const std::string& infilename) {
    std::ifstream f(infilename);
    if (!f)
        fail("cannot open " + infilename);
    ReadHstHistory(f);
}

void ChSystemDem::ReadCheckpointFile(const std::string& infilename, bool overwrite) {
    if (m_sys->initialized)
        fail("ReadCheckpointFile after Initialize");
    std::ifstream f(infilename);
    if (!f)
        fail("cannot open checkpoint " + infilename);
    std::string line;
    if (!std::getline(f, line))
        fail("empty checkpoint");
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    // "ChSystemGpu" is the header the module wrote before it was renamed (data/testing/dem/pyramid_checkpoint.dat:1)
    if (line != "ChSystemDem" && line != "ChSystemGpu" && line != "ChSystemDemMesh" && line != "ChSystemGpuMesh")
        fail("not a Chrono::Dem checkpoint: " + line);
    const unsigned int n = ReadDatParams(f, overwrite);
    while (std::getline(f, line)) {
        if (line.find("CsvParticles") != std::string::npos)
            ReadCsvParticles(f, n);
        else if (line.find("HstHistory") != std::string::npos)
            ReadHstHistory(f, n);
    }
    if (m_sys->n() != n)
        fail("checkpoint: nSpheres does not match the particle block");
    m_sys->defragment = false;  // ids stay stable (ChSystemDem.cpp:785-786); ours always are
}

// =====================================================================================================================
// ChSystemDemMesh (reference: src/chrono_dem/physics/ChSystemDem.cpp:470-655, 1274-1301, 1510-1766)
// =====================================================================================================================
ChSystemDemMesh::ChSystemDemMesh(float sphere_rad, float density, const ChVector3f& boxDims, ChVector3f O)
    : ChSystemDem(sphere_rad, density, boxDims, O) {}

ChSystemDemMesh::ChSystemDemMesh(const std::string& checkpoint) : ChSystemDem() {
    m_sys = new ChSystemDem_impl();
    m_sys->bcs.resize(NUM_RESERVED_BC_IDS);
    ReadCheckpointFile(checkpoint, true);
}

ChSystemDemMesh::~ChSystemDemMesh() {}

unsigned int ChSystemDemMesh::AddMesh(std::shared_ptr<ChTriangleMeshConnected> mesh, float mass) {
    if (m_sys->initialized)
        fail("AddMesh must be called before Initialize");
    if (!mesh)
        fail("AddMesh: null mesh");
    const unsigned int id = (unsigned int)m_meshes.size();
    m_meshes.push_back(mesh);
    m_mesh_masses.push_back(mass);
    return id;
}

unsigned int ChSystemDemMesh::AddMesh(const std::string& filename, const ChVector3f& translation,
                                      const ChMatrix33<float>& rotscale, float mass) {
    auto mesh = std::make_shared<ChTriangleMeshConnected>();
    if (!mesh->LoadWavefrontMesh(filename, true, false))
        fail("mesh " + filename + " failed to load");
    if (mesh->GetNumTriangles() == 0)
        printf("WARNING: Mesh %s has no triangles!\n", filename.c_str());
    ChMatrix33<double> rs;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            rs(i, j) = rotscale(i, j);
    mesh->Transform(ChVector3d(translation), rs);
    return AddMesh(mesh, mass);
}

std::vector<unsigned int> ChSystemDemMesh::AddMeshes(const std::vector<std::string>& objfilenames,
                                                     const std::vector<ChVector3f>& translations,
                                                     const std::vector<ChMatrix33<float>>& rotscales,
                                                     const std::vector<float>& masses) {
    const size_t size = objfilenames.size();
    if (size != rotscales.size() || size != translations.size() || size != masses.size())
        fail("mesh loading vectors must all have the same size");
    std::vector<unsigned int> ids(size);
    for (size_t i = 0; i < size; i++)
        ids[i] = AddMesh(objfilenames[i], translations[i], rotscales[i], masses[i]);
    return ids;
}

// Flatten the cached meshes into the staging soup (reference SetMeshes, ChSystemDem.cpp:543-625).
void ChSystemDemMesh::SetMeshes() {
    ChSystemDem_impl& S = *m_sys;
    std::vector<ChSystemDem_impl::MeshStage> staged(m_meshes.size());
    for (size_t m = 0; m < m_meshes.size(); m++) {
        auto& M = staged[m];
        if (m < S.meshes.size()) {  // keep a motion applied before Initialize
            M = S.meshes[m];
            M.verts9.clear();
        }
        M.mass = m_mesh_masses[m] > 0 ? (double)m_mesh_masses[m] : 1e30;
        const unsigned nt = m_meshes[m]->GetNumTriangles();
        M.verts9.reserve((size_t)nt * (m_two_sided ? 18 : 9));
        for (unsigned i = 0; i < nt; i++) {
            const ChTriangle t = m_meshes[m]->GetTriangle(i);
            const ChVector3d* v[3] = {&t.p1, &t.p2, &t.p3};
            for (int k = 0; k < 3; k++)
                for (unsigned c = 0; c < 3; c++)
                    M.verts9.push_back((*v[k])[c]);
            if (m_two_sided) {
                const ChVector3d* w[3] = {&t.p1, &t.p3, &t.p2};
                for (int k = 0; k < 3; k++)
                    for (unsigned c = 0; c < 3; c++)
                        M.verts9.push_back((*w[k])[c]);
            }
        }
        if (mesh_verbosity == CHDEM_MESH_VERBOSITY::INFO)
            printf("mesh %zu: %u facets (%s)\n", m, nt, m_two_sided ? "two-sided" : "one-sided");
    }
    S.meshes.swap(staged);
}

void ChSystemDemMesh::EnableMeshCollision(bool val) {
    m_sys->mesh_collision = val;
    if (m_sys->initialized)
        m_sys->check(dem_b200_enable_mesh_collision(m_sys->h, val ? 1 : 0), "EnableMeshCollision");
}

void ChSystemDemMesh::ApplyMeshMotion(unsigned int mesh_id, const ChVector3d& pos, const ChQuaternion<>& rot,
                                      const ChVector3d& lin_vel, const ChVector3d& ang_vel) {
    ChSystemDem_impl& S = *m_sys;
    if (mesh_id >= m_meshes.size())
        fail("ApplyMeshMotion: bad mesh id");
    const double p[3] = {pos.x(), pos.y(), pos.z()};
    const double q[4] = {rot.e0(), rot.e1(), rot.e2(), rot.e3()};
    const double v[3] = {lin_vel.x(), lin_vel.y(), lin_vel.z()};
    const double w[3] = {ang_vel.x(), ang_vel.y(), ang_vel.z()};
    if (!S.initialized) {
        if (S.meshes.size() < m_meshes.size())
            S.meshes.resize(m_meshes.size());
        auto& M = S.meshes[mesh_id];
        M.has_motion = true;
        for (int k = 0; k < 3; k++) { M.pos[k] = p[k]; M.lin[k] = v[k]; M.ang[k] = w[k]; }
        for (int k = 0; k < 4; k++) M.rot[k] = q[k];
        return;
    }
    auto& M = S.meshes[mesh_id];
    for (int k = 0; k < 3; k++) M.pos[k] = p[k];
    for (int k = 0; k < 4; k++) M.rot[k] = q[k];
    S.check(dem_b200_set_mesh_motion(S.h, (int)mesh_id, p, q, v, w), "ApplyMeshMotion");
}

unsigned int ChSystemDemMesh::GetNumMeshes() const { return (unsigned int)m_meshes.size(); }

void ChSystemDemMesh::SetStaticFrictionCoeff_SPH2MESH(float mu) { m_sys->mu[2] = mu; }
void ChSystemDemMesh::SetRollingCoeff_SPH2MESH(float mu) { m_sys->mu_roll[2] = mu; }
void ChSystemDemMesh::SetSpinningCoeff_SPH2MESH(float mu) { m_sys->mu_spin[2] = mu; }
void ChSystemDemMesh::SetKn_SPH2MESH(double v) { m_sys->Kn[2] = v; }
void ChSystemDemMesh::SetGn_SPH2MESH(double v) { m_sys->Gn[2] = v; }
void ChSystemDemMesh::SetKt_SPH2MESH(double v) { m_sys->Kt[2] = v; }
void ChSystemDemMesh::SetGt_SPH2MESH(double v) { m_sys->Gt[2] = v; }
void ChSystemDemMesh::UseMaterialBasedModel(bool val) { m_sys->use_mat_based = val; }
void ChSystemDemMesh::SetYoungModulus_MESH(double v) { m_sys->young[2] = v; }
void ChSystemDemMesh::SetPoissonRatio_MESH(double v) { m_sys->poisson[2] = v; }
void ChSystemDemMesh::SetRestitution_MESH(double v) { m_sys->cor[2] = v; }
void ChSystemDemMesh::SetAdhesionRatio_SPH2MESH(float v) { m_sys->adhesion_over_g[2] = v; }

void ChSystemDemMesh::Initialize() {
    if (!m_meshes.empty())
        SetMeshes();
    ChSystemDem::Initialize();  // the base uploads the staged meshes ahead of the spheres
}

// The reference can add the triangles to a system whose spheres are already initialised; here the facets own shape ids
// below the spheres', so they must be known at Initialize.
void ChSystemDemMesh::InitializeMeshes() {
    if (m_sys->initialized) {
        if (m_sys->num_triangles() == 0 && !m_meshes.empty())
            fail("InitializeMeshes after Initialize is not supported: add the meshes before Initialize");
        return;
    }
    if (!m_meshes.empty())
        SetMeshes();
}

double ChSystemDemMesh::AdvanceSimulation(float duration) { return ChSystemDem::AdvanceSimulation(duration); }

void ChSystemDemMesh::CollectMeshContactForces(int mesh, ChVector3d& force, ChVector3d& torque) {
    double f[3], t[3];
    m_sys->check(dem_b200_mesh_wrench(m_sys->h, mesh, f, t), "CollectMeshContactForces");
    force = ChVector3d(f[0], f[1], f[2]);
    torque = ChVector3d(t[0], t[1], t[2]);
}

void ChSystemDemMesh::CollectMeshContactForces(std::vector<ChVector3d>& forces, std::vector<ChVector3d>& torques) {
    forces.resize(m_meshes.size());
    torques.resize(m_meshes.size());
    for (size_t i = 0; i < m_meshes.size(); i++)
        CollectMeshContactForces((int)i, forces[i], torques[i]);
}

void ChSystemDemMesh::WriteCheckpointMeshParams(std::ofstream& cp) const {
    const ChSystemDem_impl& S = *m_sys;
    std::ostringstream p;
    p << "adhesionOverG_s2m: " << (float)S.adhesion_over_g[2] << "\n";
    p << "K_n_s2m: " << S.Kn[2] << "\n" << "K_t_s2m: " << S.Kt[2] << "\n";
    p << "G_n_s2m: " << S.Gn[2] << "\n" << "G_t_s2m: " << S.Gt[2] << "\n";
    p << "RollingCoeff_s2m: " << S.mu_roll[2] << "\n" << "SpinningCoeff_s2m: " << S.mu_spin[2] << "\n";
    p << "StaticFrictionCoeff_s2m: " << S.mu[2] << "\n";
    p << "MeshCollisionEnabled: " << (S.mesh_collision ? 1 : 0) << "\n";
    cp << p.str();
}

void ChSystemDemMesh::WriteCheckpointFile(const std::string& outfilename) {
    std::ofstream cp(outfilename, std::ios::out);
    cp << "ChSystemDemMesh\n";
    WriteCheckpointParams(cp);
    WriteCheckpointMeshParams(cp);
    cp << "ParamsEnd\n\n";
    cp << "CsvParticles\n";
    const unsigned int flags = m_sys->out_flags;
    m_sys->out_flags = VEL_COMPONENTS | FIXITY | ANG_VEL_COMPONENTS;
    WriteCsvParticles(cp);
    m_sys->out_flags = flags;
    cp << "\n";
    if (m_sys->friction != CHDEM_FRICTION_MODE::FRICTIONLESS) {
        cp << "HstHistory\n";
        WriteHstHistory(cp);
        cp << "\n";
    }
}

void ChSystemDemMesh::ReadCheckpointFile(const std::string& infilename, bool overwrite) {
    ChSystemDem::ReadCheckpointFile(infilename, overwrite);  // SetParamsFromIdentifier is virtual: mesh keys are ours
}

bool ChSystemDemMesh::SetParamsFromIdentifier(const std::string& id, std::istringstream& iss, bool overwrite) {
    ChSystemDem_impl& S = *m_sys;
    unsigned u;
    if (id == "adhesionOverG_s2m") iss >> S.adhesion_over_g[2];
    else if (id == "K_n_s2m") iss >> S.Kn[2];
    else if (id == "K_t_s2m") iss >> S.Kt[2];
    else if (id == "G_n_s2m") iss >> S.Gn[2];
    else if (id == "G_t_s2m") iss >> S.Gt[2];
    else if (id == "RollingCoeff_s2m") iss >> S.mu_roll[2];
    else if (id == "SpinningCoeff_s2m") iss >> S.mu_spin[2];
    else if (id == "StaticFrictionCoeff_s2m") iss >> S.mu[2];
    else if (id == "MeshCollisionEnabled") { iss >> u; S.mesh_collision = u != 0; }
    else return ChSystemDem::SetParamsFromIdentifier(id, iss, overwrite);
    return true;
}

// ASCII VTK unstructured grid of mesh i at its current frame (reference: ChSystemDem.cpp:1564-1613).
static void write_vtk(std::ostream& o, const std::vector<const ChTriangleMeshConnected*>& ms,
                      const std::vector<const ChSystemDem_impl::MeshStage*>& frames) {
    size_t nv = 0, nf = 0;
    for (auto* m : ms) { nv += m->GetCoordsVertices().size(); nf += m->GetIndicesVertices().size(); }
    o << "# vtk DataFile Version 2.0\nVTK from simulation\nASCII\n\n\nDATASET UNSTRUCTURED_GRID\n";
    o << "POINTS " << nv << " float" << std::endl;
    for (size_t k = 0; k < ms.size(); k++) {
        const auto* F = frames[k];
        ChQuaternion<double> q(1, 0, 0, 0);
        double p[3] = {0, 0, 0};
        if (F) { q = ChQuaternion<double>(F->rot[0], F->rot[1], F->rot[2], F->rot[3]); p[0] = F->pos[0]; p[1] = F->pos[1]; p[2] = F->pos[2]; }
        ChMatrix33<double> A(q);
        for (auto& v : ms[k]->GetCoordsVertices()) {
            const ChVector3d w = A * v;
            o << (float)(w.x() + p[0]) << " " << (float)(w.y() + p[1]) << " " << (float)(w.z() + p[2]) << std::endl;
        }
    }
    o << "\n\nCELLS " << nf << " " << 4 * nf << std::endl;
    size_t off = 0;
    for (auto* m : ms) {
        for (auto& f : m->GetIndicesVertices())
            o << "3 " << f.x() + (int)off << " " << f.y() + (int)off << " " << f.z() + (int)off << std::endl;
        off += m->GetCoordsVertices().size();
    }
    o << "\n\nCELL_TYPES " << nf << std::endl;
    for (size_t j = 0; j < nf; j++)
        o << "5 " << std::endl;
}

static std::string vtk_name(const std::string& n) {
    const std::string tail = n.substr(n.length() - std::min(n.length(), (size_t)4));
    return (tail == ".vtk" || tail == ".VTK") ? n : n + ".vtk";
}

void ChSystemDemMesh::WriteMesh(const std::string& outfilename, unsigned int i) const {
    if (m_sys->out_mode == CHDEM_OUTPUT_MODE::NONE)
        return;
    if (i >= m_meshes.size()) {
        printf("WARNING: attempted to write mesh %u, yet only %zu meshes present. No mesh file generated.\n", i, m_meshes.size());
        return;
    }
    std::ofstream f(vtk_name(outfilename), std::ios::out);
    write_vtk(f, {m_meshes[i].get()}, {i < m_sys->meshes.size() ? &m_sys->meshes[i] : nullptr});
}

void ChSystemDemMesh::WriteMeshes(const std::string& outfilename) const {
    if (m_sys->out_mode == CHDEM_OUTPUT_MODE::NONE)
        return;
    if (m_meshes.empty()) {
        printf("WARNING: attempted to write meshes to file yet no mesh found in system cache. No mesh file generated.\n");
        return;
    }
    std::vector<const ChTriangleMeshConnected*> ms;
    std::vector<const ChSystemDem_impl::MeshStage*> fr;
    for (size_t i = 0; i < m_meshes.size(); i++) {
        ms.push_back(m_meshes[i].get());
        fr.push_back(i < m_sys->meshes.size() ? &m_sys->meshes[i] : nullptr);
    }
    std::ofstream f(vtk_name(outfilename), std::ios::out);
    write_vtk(f, ms, fr);
}

}  // namespace dem
}  // namespace chrono

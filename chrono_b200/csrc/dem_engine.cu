// =============================================================================
// dem_engine.cu -- host side of the B200 SMC-DEM engine and its C ABI (include/chrono_b200_dem.h).
//
// Replaces, for the hot path, ChSystemDem_impl::initializeSpheres / AdvanceSimulation / getters
// (reference: src/chrono_dem/physics/ChSystemDem_impl.cpp:1130-1166, src/chrono_dem/gpu/ChDemSMC.cu:619-691).
// Unlike the reference there is no managed memory, no device synchronisation between kernels and no host
// round trip inside a step: the launches of a step are enqueued on one stream and replayed from a CUDA graph.
// Every per-step decision (rebuild the Verlet candidate lists or not, which ping-pong buffer is live) is taken on
// the device (Ctrl block, k_step_begin), so the same one-step graph is valid for every step.
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/chrono_b200_dem.h"
#include "dem_kernels.cuh"

using namespace demb200;

#ifndef DEMB200_COND_GRAPH
#define DEMB200_COND_GRAPH 1
#endif
#ifndef DEMB200_TILE_DEFAULT
#define DEMB200_TILE_DEFAULT 0
#endif

namespace {
thread_local std::string g_create_error;
}

struct dem_b200_system {
    dem_b200_config cfg{};
    Params P{};
    Buffers B{};
    WallSet W{};                       // host copy of the walls (device copy: B.walls)
    MeshSet M{};                       // host copy of the mesh bodies (device copy: B.meshes)
    std::vector<double> h_tri;         // triangle soup, body frame, 9 doubles per triangle
    std::vector<uint32_t> h_tri_mesh;  // owner mesh of each triangle
    double mesh_radius[kMaxMeshes] = {0};  // largest vertex distance from the body origin (bound of a rotation's travel)
    MeshBody* h_mesh_pin = nullptr;    // pinned staging of MeshSet::m, a ring of kMeshRing copies (co-simulation applies mesh
    unsigned mesh_ring_pos = 0;        // motion every step: no host wait except when the ring wraps)
    bool cls_override[3] = {false, false, false};
    dem_b200_contact_class cls[3]{};   // explicit contact-class coefficients (dem_b200_set_contact_class)
    // scene staged on the host until initialize()
    std::vector<double> h_pos, h_vel, h_om, h_rad;
    std::vector<uint8_t> h_fixed;
    bool any_fixed = false;
    bool initialized = false;
    cudaStream_t stream = nullptr;
    cudaStream_t cap_stream = nullptr;  // graph capture only
    cudaStream_t cap_stream2 = nullptr; // body of the conditional node
    cudaStream_t side_stream = nullptr; // slab mode: halo pick-up + second force pass next to the first pass
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    unsigned ntiles = 0;               // scan tiles covering the search-cell capacity
    cudaGraphExec_t graph1 = nullptr;  // one step
    bool recording = false;
    bool user_recording = false;       // dem_b200_enable_recording asked for by the caller (contact info piggybacks on the same kernels)
    bool contact_info = false;
    double* d_cinfo_out = nullptr;
    size_t max_pairs = 0;
    bool use_hrel = false;
    bool track_wall_forces = false;
    // slab decomposition (one process per GPU)
    bool mgpu = false;
    bool own_stream = true;
    size_t mg_cap = 0;                 // capacity (local spheres incl. ghosts)
    double mg_rmax = 0;                // global largest radius
    std::vector<uint32_t> h_ids;       // global stable ids of the spheres given to set_spheres
    unsigned mg_n_own = 0, mg_n_local = 0;
    unsigned mg_ns[2] = {0, 0};        // owned spheres sent as ghosts to the left / right neighbour
    unsigned mg_ng[2] = {0, 0};        // ghosts received from the left / right neighbour
    int mg_phase = 0;                  // 0 idle, 1 extracted, 2 ghosts selected
    bool mg_remap_pending = false;
    unsigned* d_count = nullptr;
    // direct P2P halo (dem_b200_p2p_*)
    bool p2p = false;
    P2PDev X{};
    void* p2p_region = nullptr;          // my exposed region (control block + landing buffers)
    void* p2p_mapped[kMaxRanks] = {nullptr};  // peers' regions as opened here
    size_t p2p_records = 0;
    unsigned long long* h_vote = nullptr;     // pinned, mapped
    unsigned long long step_no = 0;      // host mirror of Ctrl::nsteps (steps enqueued so far)
    unsigned long long p2p_rebuild_seq = 0;
    unsigned long long slab_rebuilds = 0;  // host-driven slab rebuilds (dem_b200_mgpu_finish_rebuild)
    unsigned long long mg_last_rebuild_step = 0, mg_last_interval = 0;  // time steps between the last two slab rebuilds
    size_t owned_export_n = (size_t)-1;    // dem_b200_export_owned bookkeeping for dem_b200_import_owned
    unsigned long long owned_export_step = ~0ull, owned_export_seq = ~0ull;
    int one_step_calls = 0;
    double time = 0.0;
    std::string err;
    // scratch (device, by user index) and pinned host staging
    double* d_pos3 = nullptr; double* d_vel3 = nullptr; double* d_om3 = nullptr;
    double* d_red = nullptr;            // reduction scratch (2 x 8 bytes)
    double* h_pin = nullptr;            // pinned, 16 doubles
    bool export_valid = false;
    double* d_acc3 = nullptr;        // user-order acceleration of the last step (allocated at the first dem_b200_get_accel)
    bool accel_src_valid = false;    // the last thing that changed the state was a step (the other ping-pong buffer is its input)
    bool accel_export_valid = false;
    std::vector<void*> allocs;
    // history staged before initialize (add_history)
    struct HRow { uint32_t owner, other; double d[3], dur, rel; };
    std::vector<HRow> h_hist;
};

constexpr unsigned kMeshRing = 16;

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            return DEMB200_ECUDA;                                                                \
        }                                                                                        \
    } while (0)

namespace {

template <class T>
int dev_alloc(dem_b200_system* s, T** p, size_t count) {
    void* q = nullptr;
    CU(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    s->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

// ChContactMaterialCompositeSMC (src/chrono/physics/ChContactMaterialSMC.cpp:107-130): float arithmetic, default
// composition strategy (src/chrono/physics/ChContactMaterial.h:161-169).
Comp make_comp(const dem_b200_material& m1, const dem_b200_material& m2) {
    float inv_E = (1 - m1.poisson * m1.poisson) / m1.young + (1 - m2.poisson * m2.poisson) / m2.young;
    float inv_G = 2 * (2 - m1.poisson) * (1 + m1.poisson) / m1.young + 2 * (2 - m2.poisson) * (1 + m2.poisson) / m2.young;
    float E_eff = 1 / inv_E;
    float G_eff = 1 / inv_G;
    Comp c;
    c.E_eff = E_eff;
    c.G_eff = G_eff;
    c.mu = std::min<float>(m1.mu_s, m2.mu_s);
    c.mu_roll = std::min<float>(m1.mu_roll, m2.mu_roll);
    c.mu_spin = std::min<float>(m1.mu_spin, m2.mu_spin);
    c.cr = std::min<float>(m1.cr, m2.cr);
    c.adh = std::min<float>(m1.adhesion, m2.adhesion);
    c.adh_dmt = std::min<float>(m1.adhesion_dmt, m2.adhesion_dmt);
    c.adh_perko = std::min<float>(m1.adhesion_perko, m2.adhesion_perko);
    c.kn = (m1.kn + m2.kn) / 2;
    c.kt = (m1.kt + m2.kt) / 2;
    c.gn = (m1.gn + m2.gn) / 2;
    c.gt = (m1.gt + m2.gt) / 2;
    // material-only factor of the Hertz damping, ChIterativeSolverMulticoreSMC.cpp:283-287
    const double eps = 2.220446049250313e-16, kPI = 3.141592653589793238462643383279;
    double loge = (c.cr < eps) ? std::log(eps) : std::log(c.cr);
    double beta = loge / std::sqrt(loge * loge + kPI * kPI);
    c.hertz_damp = -2 * std::sqrt(5.0 / 6) * beta;
    c.gt_ratio = std::sqrt(4.0 * c.G_eff / c.E_eff);  // sqrt(St / Sn), :279-280
    return c;
}

// AbsRotate, src/chrono/multicore_math/real4.cpp:168-187
void abs_rotate(const double q[4], const double v[3], double out[3]) {
    double e0e0 = q[0] * q[0], e1e1 = q[1] * q[1], e2e2 = q[2] * q[2], e3e3 = q[3] * q[3];
    double e0e1 = q[0] * q[1], e0e2 = q[0] * q[2], e0e3 = q[0] * q[3];
    double e1e2 = q[1] * q[2], e1e3 = q[1] * q[3], e2e3 = q[2] * q[3];
    out[0] = std::abs((e0e0 + e1e1) * 2 - 1) * v[0] + std::abs((e1e2 - e0e3) * 2) * v[1] + std::abs((e1e3 + e0e2) * 2) * v[2];
    out[1] = std::abs((e1e2 + e0e3) * 2) * v[0] + std::abs((e0e0 + e2e2) * 2 - 1) * v[1] + std::abs((e2e3 - e0e1) * 2) * v[2];
    out[2] = std::abs((e1e3 - e0e2) * 2) * v[0] + std::abs((e2e3 + e0e1) * 2) * v[1] + std::abs((e0e0 + e3e3) * 2 - 1) * v[2];
}

void refresh_params(dem_b200_system* s) {
    const dem_b200_config& c = s->cfg;
    Params& P = s->P;
    P.K = c.history_slots > 0 ? std::min(c.history_slots, kMaxSlots) : 16;
    P.Kn = c.neighbor_slots > 0 ? std::min(c.neighbor_slots, kMaxNeighbors) : 32;
    P.force_model = c.force_model;
    P.adhesion_model = c.adhesion_model;
    P.tang_mode = c.tangential_mode;
    P.use_mat_props = c.use_mat_props;
    P.integrator = c.integrator;
    P.char_vel = c.char_vel;
    P.min_slip = c.min_slip_vel;
    P.min_roll = c.min_roll_vel;
    P.min_spin = c.min_spin_vel;
    P.dt = c.dt;
    for (int k = 0; k < 3; k++) {
        P.g[k] = c.gravity[k];
        P.bins[k] = c.bins_per_axis[k];
    }
    P.mass_coef = c.mass_coef;
    P.wall_mass = c.wall_mass;
    P.comp[0] = make_comp(c.material[DEMB200_MAT_SPHERE], c.material[DEMB200_MAT_SPHERE]);
    P.comp[1] = make_comp(c.material[DEMB200_MAT_WALL], c.material[DEMB200_MAT_SPHERE]);
    P.comp[2] = make_comp(c.material[DEMB200_MAT_MESH], c.material[DEMB200_MAT_SPHERE]);
    for (int k = 0; k < 3; k++)
        if (s->cls_override[k]) {
            const dem_b200_contact_class& o = s->cls[k];
            Comp& cm = P.comp[k];
            cm.E_eff = o.E_eff; cm.G_eff = o.G_eff; cm.mu = o.mu; cm.mu_roll = o.mu_roll; cm.mu_spin = o.mu_spin;
            cm.cr = o.cr; cm.adh = o.adhesion; cm.kn = o.kn; cm.kt = o.kt; cm.gn = o.gn; cm.gt = o.gt;
            const double eps = 2.220446049250313e-16, kPI = 3.141592653589793238462643383279;
            double loge = (cm.cr < eps) ? std::log(eps) : std::log(cm.cr);
            cm.hertz_damp = -2 * std::sqrt(5.0 / 6) * (loge / std::sqrt(loge * loge + kPI * kPI));
            cm.gt_ratio = (cm.E_eff > 0) ? std::sqrt(4.0 * cm.G_eff / cm.E_eff) : 0.0;
        }
    // union of the box-wall AABBs
    WallSet& WS = s->W;
    WS.has_bb = 0;
    for (int w = 0; w < P.nW; w++) {
        Wall& W = WS.w[w];
        if (W.type != WALL_BOX && !(W.type == WALL_SPHERE && W.hdims[1] > 0))
            continue;
        double ext[3];
        if (W.type == WALL_BOX)
            abs_rotate(W.rot, W.hdims, ext);  // ComputeAABBBox, ChCollisionSystemMulticore.cpp:395-406 (envelope 0)
        else
            ext[0] = ext[1] = ext[2] = W.hdims[0];  // a fixed sphere shape: ComputeAABBSphere, :376-385
        for (int k = 0; k < 3; k++) {
            W.amin[k] = W.pos[k] - ext[k];
            W.amax[k] = W.pos[k] + ext[k];
            if (!WS.has_bb) {
                WS.bb_min[k] = W.amin[k];
                WS.bb_max[k] = W.amax[k];
            } else {
                WS.bb_min[k] = std::min(WS.bb_min[k], W.amin[k]);
                WS.bb_max[k] = std::max(WS.bb_max[k], W.amax[k]);
            }
        }
        WS.has_bb = 1;
    }
    P.track_wall_forces = s->track_wall_forces ? 1 : 0;
    P.external_rebuild = s->mgpu ? 1 : 0;
    P.nT = (unsigned)s->h_tri_mesh.size();
    P.shape_base = (unsigned)P.nW + P.nT;  // Multicore numbering: wall shapes, then mesh triangles, then spheres (Q12)
    // Verlet skin: negative -> default 0.25 * largest radius (set at initialize, when radii are known)
    if (c.verlet_skin >= 0)
        P.skin = c.verlet_skin;
    else
        P.skin = (s->mgpu ? 0.35 : 0.25) * P.rmax;  // slab mode: rebuilds cost a neighbour exchange and the skin is not adaptive
    P.skin_adaptive = (c.verlet_skin < 0 && !s->mgpu) ? 1 : 0;
    P.skin_max = std::max(P.skin, 0.5 * P.rmax);
    P.skin_tri = std::max(P.skin_max, 1.5 * P.rmax);
}

bool need_roll(const dem_b200_system* s) {
    for (int k = 0; k < (s->P.nT ? 3 : 2); k++)
        if (s->P.comp[k].mu_roll > 0 || s->P.comp[k].mu_spin > 0)
            return true;
    return false;
}

bool fast_path(const dem_b200_system* s) {
    return s->P.force_model == DEMB200_HERTZ && s->P.adhesion_model == DEMB200_ADH_CONSTANT;
}

// The tile kernel (shared-memory staging of the neighbour bins): Hertz fast paths, no meshes, no recording.
template <bool H, bool R, int F>
void launch_tile(dem_b200_system* s, const Buffers& B, unsigned pass) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_force_tile<H, R, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes);
        attr_set = true;
    }
    const unsigned tiles = s->P.cell_cap / (unsigned)kTileCells + 1u;  // upper bound of the tile count; surplus blocks leave at once
    k_force_tile<H, R, F><<<tiles, kTileThreads, kTileSmemBytes, s->stream>>>(s->P, B, pass);
}

bool use_tile_kernel(const dem_b200_system* s) {
    return s->P.tiled && !s->P.nT && fast_path(s) && !s->recording;
}

// Shared-memory carve-out of the force kernel: just what its resident blocks need (16 one-warp blocks x (2.5 KB contact list + 1 KB
// reserved) -> the 64 KB configuration), so that the rest of the 256 KB stays L1 for the partner gathers; the driver's default keeps
// the 100 KB configuration (r02dyn: 281.2 -> 277.4 us).  DEMB200_CARVEOUT=<percent> overrides (tuning experiments).
template <bool H, bool R, int F, bool REC, bool MESH>
void force_kernel_attrs(const Params& P) {
    static int last = -1;
    const int blocks = R ? DEMB200_ROLL_MINBLOCKS : DEMB200_FORCE_MINBLOCKS;
    int pct = (int)((100.0 * blocks * (double)(force_smem_bytes(P, H) + 1024u)) / (228.0 * 1024.0) + 0.999);
    if (const char* e = getenv("DEMB200_CARVEOUT"))
        pct = atoi(e);
    pct = pct > 100 ? 100 : pct;
    if (pct == last)
        return;
    last = pct;
    cudaFuncSetAttribute(k_force_integrate<H, R, F, REC, MESH>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// pass 0: all spheres; 1 / 2: the spheres without / with a ghost among their candidates (slab mode, direct halo: see enqueue_post)
template <bool REC>
void launch_force(dem_b200_system* s, const Buffers& B, unsigned blocks, unsigned pass = 0u) {
    const Params& P = s->P;
    const bool hist = (P.tang_mode == DEMB200_TANG_MULTISTEP);
    const bool roll = need_roll(s);
    const int fast = fast_path(s) ? (P.use_mat_props ? 1 : 2) : 0;
    if (!REC && use_tile_kernel(s)) {
        const int tsel = (hist ? 4 : 0) + (roll ? 2 : 0) + (fast - 1);
        switch (tsel) {
            case 0: launch_tile<false, false, 1>(s, B, pass); break;
            case 1: launch_tile<false, false, 2>(s, B, pass); break;
            case 2: launch_tile<false, true, 1>(s, B, pass); break;
            case 3: launch_tile<false, true, 2>(s, B, pass); break;
            case 4: launch_tile<true, false, 1>(s, B, pass); break;
            case 5: launch_tile<true, false, 2>(s, B, pass); break;
            case 6: launch_tile<true, true, 1>(s, B, pass); break;
            default: launch_tile<true, true, 2>(s, B, pass); break;
        }
        return;
    }
    const int sel = (hist ? 6 : 0) + (roll ? 3 : 0) + fast;
#define LF(H, R, F)                                                                                  \
    do {                                                                                             \
        if (P.nT)                                                                                    \
            force_kernel_attrs<H, R, F, REC, true>(P);                                               \
        else                                                                                         \
            force_kernel_attrs<H, R, F, REC, false>(P);                                              \
        if (P.nT)                                                                                    \
            k_force_integrate<H, R, F, REC, true><<<blocks, kForceThreads, force_smem_bytes(P, H), s->stream>>>(P, B, pass);    \
        else                                                                                         \
            k_force_integrate<H, R, F, REC, false><<<blocks, kForceThreads, force_smem_bytes(P, H), s->stream>>>(P, B, pass);   \
    } while (0)
    switch (sel) {
        case 0: LF(false, false, 0); break;
        case 1: LF(false, false, 1); break;
        case 2: LF(false, false, 2); break;
        case 3: LF(false, true, 0); break;
        case 4: LF(false, true, 1); break;
        case 5: LF(false, true, 2); break;
        case 6: LF(true, false, 0); break;
        case 7: LF(true, false, 1); break;
        case 8: LF(true, false, 2); break;
        case 9: LF(true, true, 0); break;
        case 10: LF(true, true, 1); break;
        default: LF(true, true, 2); break;
    }
#undef LF
}

constexpr int kNumKernels = 9;
const char* kKernelNames[kNumKernels] = {"k_step_begin", "k_bin_count", "k_scan_tile_sums", "k_scan_sums",
                                         "k_scan_apply", "k_scatter_perm", "k_gather_sorted", "k_build_list",
                                         "k_force_integrate"};

// Enqueue one step.  ev: optional kNumKernels+1 events recorded around each launch (profiling).
// One step = halo + control kernel | rebuild kernels (they return at once unless k_step_begin said "rebuild") | force kernel.
// The three parts are separate so that the step graph can put the middle one into a conditional node.
void enqueue_pre(dem_b200_system* s, cudaStream_t st, unsigned long long cond) {
    k_step_begin<<<1, 32, 0, st>>>(s->P, s->B, cond);
}

// Direct halo of slab mode (dem_b200_p2p_*): at the END of a step the boundary spheres' new state goes straight into the
// neighbours' landing buffers (NVLink stores, tagged with the number of the step that will consume it) and the rebuild vote is
// cast; the consumer picks the data up on a side stream next to the first pass of its force kernel (enqueue_post).
void enqueue_halo_pack(dem_b200_system* s, cudaStream_t st) {
    for (int d = 0; d < 2; d++)
        if (s->mg_ns[d])
            k_p2p_pack<<<(s->mg_ns[d] + 255) / 256, 256, 0, st>>>(s->B, s->X, d, s->mg_ns[d]);
}
void enqueue_vote(dem_b200_system* s, cudaStream_t st) { k_p2p_vote<<<1, 32, 0, st>>>(s->P, s->B, s->X); }

// ev: optional events recorded after each of the seven rebuild kernels (profiling), starting at ev[*k]
void enqueue_rebuild(dem_b200_system* s, cudaStream_t st, cudaEvent_t* ev, int* k) {
    const Params& P = s->P;
    Buffers& B = s->B;
    const unsigned N = P.N;
    const unsigned nb256 = (N + 255) / 256;
    auto mark = [&]() {
        if (ev)
            cudaEventRecord(ev[(*k)++], st);
    };
    k_bin_count<<<nb256, 256, 0, st>>>(P, B);
    mark();
    k_scan_tile_sums<<<s->ntiles, kScanThreads, 0, st>>>(B, 0);
    mark();
    k_scan_sums<<<1, kScanThreads, 0, st>>>(B, 0);
    mark();
    k_scan_apply<<<s->ntiles, kScanThreads, 0, st>>>(P, B, 0);
    mark();
    k_scatter_perm<<<nb256, 256, 0, st>>>(P, B);
    mark();
    k_gather_sorted<<<nb256, 256, 0, st>>>(P, B);
    mark();
    if (P.nT) {  // mesh triangles -> search cells (timed with k_build_list in the profile)
        const unsigned tb = (P.nT + 255) / 256;
        k_tri_count<<<tb, 256, 0, st>>>(P, B);
        k_scan_tile_sums<<<s->ntiles, kScanThreads, 0, st>>>(B, 1);
        k_scan_sums<<<1, kScanThreads, 0, st>>>(B, 1);
        k_scan_apply<<<s->ntiles, kScanThreads, 0, st>>>(P, B, 1);
        k_tri_fill<<<tb, 256, 0, st>>>(P, B);
    }
    k_build_list<<<(N + kListThreads - 1) / kListThreads, kListThreads, 0, st>>>(P, B);
    mark();
}

void enqueue_post(dem_b200_system* s, cudaStream_t st) {
    const Params& P = s->P;
    Buffers& B = s->B;
    const unsigned N = P.N;
    const unsigned fb = (N + kForceThreads - 1) / kForceThreads;
    cudaStream_t keep = s->stream;
    s->stream = st;  // launch_force uses s->stream
    if (s->recording) {
        k_record_bins<<<(N + 255) / 256, 256, 0, st>>>(P, B);
        launch_force<true>(s, B, fb);
    } else if (s->p2p && !s->mg_remap_pending) {
        // Two passes.  Pass 2 = the ghost senders (every owned sphere within the ghost cut of a slab face: they include every
        // sphere that has a ghost among its candidates), pass 1 = everybody else: 95 % of a 1 M-sphere slab, needs no halo.
        // Side stream, next to pass 1 (a fork / join of the step graph): pick up the halo the neighbours sent during THEIR
        // previous step -> pass 2 -> send the new state of the senders for the neighbours' next step.  Main stream: pass 1, then
        // the rebuild vote (it needs the displacement of every sphere).  Critical path: step control -> pass 1 -> vote.
        if (!s->side_stream) {
            cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
        }
        cudaEventRecord(s->ev_fork, st);
        cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0);
        launch_force<false>(s, B, fb, 1u);
        s->stream = s->side_stream;
        for (int d = 0; d < 2; d++)
            if (s->mg_ng[d])
                k_p2p_unpack<<<(s->mg_ng[d] + 255) / 256, 256, 0, s->side_stream>>>(B, s->X, d, s->mg_ng[d], 1);
        const unsigned nb = s->mg_ns[0] + s->mg_ns[1];
        if (nb)
            launch_force<false>(s, B, (nb + kForceThreads - 1) / kForceThreads, 2u);
        enqueue_halo_pack(s, s->side_stream);
        cudaEventRecord(s->ev_join, s->side_stream);
        cudaStreamWaitEvent(st, s->ev_join, 0);
        s->stream = st;
        enqueue_vote(s, st);
    } else {
        // (slab mode right after a rebuild: the ghosts that were just exchanged are current, one pass; run_steps sends the halo
        // once the sorted slots of the senders are known)
        launch_force<false>(s, B, fb);
    }
    s->stream = keep;
}

// Enqueue one step.  ev: optional kNumKernels+1 events recorded around each launch (profiling).
int enqueue_step(dem_b200_system* s, cudaEvent_t* ev) {
    cudaStream_t st = s->stream;
    int k = 0;
    if (ev)
        cudaEventRecord(ev[k++], st);
    enqueue_pre(s, st, 0ull);
    if (ev)
        cudaEventRecord(ev[k++], st);
    enqueue_rebuild(s, st, ev, &k);
    enqueue_post(s, st);
    if (ev)
        cudaEventRecord(ev[k++], st);
    CU(cudaGetLastError());
    s->time += s->P.dt;
    s->step_no++;
    s->export_valid = false;
    s->accel_src_valid = true;
    s->accel_export_valid = false;
    return 0;
}

void drop_graph(dem_b200_system* s) {
    if (s->graph1) {
        cudaGraphExecDestroy(s->graph1);
        s->graph1 = nullptr;
    }
}

// The step graph: [halo, k_step_begin] -> IF(rebuild) { rebuild kernels } -> [force kernel, vote].  k_step_begin sets the
// condition on the device (cudaGraphSetConditional), so a step that does not rebuild launches nothing in between.  Falls
// back to the flat capture (rebuild kernels that return at once) if conditional nodes are not available.
int build_graph_conditional(dem_b200_system* s, cudaGraph_t* out) {
#if DEMB200_COND_GRAPH
    cudaStream_t cs = s->cap_stream;
    if (!s->cap_stream2 && cudaStreamCreateWithFlags(&s->cap_stream2, cudaStreamNonBlocking) != cudaSuccess)
        return 1;
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        return 1;
    cudaGraph_t g = nullptr;
    auto bail = [&]() {
        cudaGraph_t junk = nullptr;
        cudaStreamEndCapture(cs, &junk);
        if (junk)
            cudaGraphDestroy(junk);
        cudaGetLastError();
        return 1;
    };
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    if (cudaStreamGetCaptureInfo_v2(cs, &status, nullptr, &g, &deps, &ndeps) != cudaSuccess || !g)
        return bail();
    cudaGraphConditionalHandle handle;
    if (cudaGraphConditionalHandleCreate(&handle, g, 0, cudaGraphCondAssignDefault) != cudaSuccess)
        return bail();
    enqueue_pre(s, cs, (unsigned long long)handle);
    if (cudaStreamGetCaptureInfo_v2(cs, &status, nullptr, &g, &deps, &ndeps) != cudaSuccess)
        return bail();
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeIf;
    cp.conditional.size = 1;
    cudaGraphNode_t cnode;
    if (cudaGraphAddNode(&cnode, g, deps, ndeps, &cp) != cudaSuccess)
        return bail();
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    if (cudaStreamBeginCaptureToGraph(s->cap_stream2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        return bail();
    enqueue_rebuild(s, s->cap_stream2, nullptr, nullptr);
    cudaGraph_t body_out = nullptr;
    if (cudaStreamEndCapture(s->cap_stream2, &body_out) != cudaSuccess)
        return bail();
    if (cudaStreamUpdateCaptureDependencies(cs, &cnode, 1, cudaStreamSetCaptureDependencies) != cudaSuccess)
        return bail();
    enqueue_post(s, cs);
    if (cudaStreamEndCapture(cs, &g) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    *out = g;
    return 0;
#else
    (void)s; (void)out;
    return 1;
#endif
}

int build_graph(dem_b200_system* s) {
    cudaGraph_t g = nullptr;
    // capture on a private stream: the caller's stream may be the legacy default stream (slab mode runs on torch's
    // current stream), which cannot be captured; the instantiated graph is launched into s->stream
    if (!s->cap_stream)
        CU(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
    // The conditional step graph is opt-in (DEMB200_COND_GRAPH=1 in the environment): it saves the seven empty rebuild
    // launches of a non-rebuilding step (32.0 vs 32.4 - 33.2 ms per 100 steps at 1 M spheres), but ncu cannot see the kernel
    // nodes of a graph that holds a conditional node, so the default is the flat capture whose launches are all profilable.
    // Slab mode re-captures the graph after every slab rebuild (the local sphere count changes): there the flat capture,
    // which is cheaper to build, wins anyway (2 x B200: 36.2 vs 37.3 ms per 100 steps).
    // The conditional step graph is the default outside slab mode (DEMB200_COND_GRAPH=0 in the environment selects the flat
    // capture): a step that does not rebuild launches k_step_begin and the force kernel only, instead of seven more kernels that
    // return at once (3 - 5 % of a 1 M-sphere step).  ncu cannot see the kernel nodes of a graph that holds a conditional node:
    // profile with DEMB200_COND_GRAPH=0 or through direct launches (scripts/profile_kernels.py does).
    // Slab mode re-captures the graph after every slab rebuild (the local sphere count changes): there the flat capture,
    // which is cheaper to build, wins anyway (2 x B200: 36.2 vs 37.3 ms per 100 steps).
    static const bool want_cond = [] {
        const char* e = getenv("DEMB200_COND_GRAPH");
        return !(e && e[0] == '0');
    }();
    // slab mode re-captures the graph after every slab rebuild: the conditional form (dearer to build, cheaper to run) only when
    // the rebuilds have been coming more than 150 steps apart (a settled bed), the flat one while the bed flows
    if (want_cond && (!s->mgpu || s->mg_last_interval >= 150ull) && build_graph_conditional(s, &g) == 0) {
        cudaError_t ei = cudaGraphInstantiate(&s->graph1, g, 0);
        cudaGraphDestroy(g);
        if (ei == cudaSuccess)
            return 0;
        cudaGetLastError();
        s->graph1 = nullptr;
    }
    g = nullptr;
    cudaStream_t run_stream = s->stream;
    s->stream = s->cap_stream;
    cudaError_t e0 = cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (e0 != cudaSuccess) {
        s->stream = run_stream;
        s->err = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e0);
        return DEMB200_ECUDA;
    }
    double t = s->time;
    const unsigned long long sn = s->step_no;
    int rc = enqueue_step(s, nullptr);
    s->time = t;  // capturing does not advance time
    s->step_no = sn;
    cudaError_t e = cudaStreamEndCapture(s->cap_stream, &g);
    s->stream = run_stream;
    if (rc)
        return rc;
    if (e != cudaSuccess) {
        s->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e);
        return DEMB200_ECUDA;
    }
    CU(cudaGraphInstantiate(&s->graph1, g, 0));
    cudaGraphDestroy(g);
    return 0;
}

int check_device_error(dem_b200_system* s) {
    unsigned e = 0;
    CU(cudaMemcpyAsync(s->h_pin, &s->B.ctrl->err, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(&e, s->h_pin, sizeof(unsigned));
    if (!e)
        return 0;
    if (e & ERR_NAN) { s->err = "non-finite sphere state"; return DEMB200_ENAN; }
    if (e & ERR_HISTORY_OVERFLOW) { s->err = "a sphere has more contacts than history_slots"; return DEMB200_EHISTORY; }
    if (e & ERR_NEIGHBOR_OVERFLOW) { s->err = "a sphere has more neighbour candidates than neighbor_slots"; return DEMB200_ENEIGHBORS; }
    if (e & ERR_PAIR_CAPACITY) { s->err = "pair recording buffer overflow"; return DEMB200_ECAPACITY; }
    if (e & ERR_SKIN_EXCEEDED) { s->err = "slab mode: the Verlet skin was used up before the driver rebuilt the lists"; return DEMB200_EINVAL; }
    if (e & ERR_MESH_CAPACITY) { s->err = "mesh triangles cover more search cells than reserved (triangles much larger than the spheres: subdivide the mesh)"; return DEMB200_ECAPACITY; }
    s->err = "unknown device error";
    return DEMB200_ECUDA;
}

int read_ctrl(dem_b200_system* s, Ctrl* out) {
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(out, s->B.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost));
    return 0;
}

int recompute_bbox(dem_b200_system* s) {
    unsigned long long init[6];
    for (int k = 0; k < 3; k++) {
        init[k] = s->W.has_bb ? enc_ord_h(s->W.bb_min[k]) : enc_ord_h(INFINITY);
        init[3 + k] = s->W.has_bb ? enc_ord_h(s->W.bb_max[k]) : enc_ord_h(-INFINITY);
    }
    unsigned long long sinit[6];
    for (int k = 0; k < 3; k++) {
        sinit[k] = enc_ord_h(INFINITY);
        sinit[3 + k] = enc_ord_h(-INFINITY);
    }
    CU(cudaMemcpyAsync(s->B.ctrl->bbox, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->B.ctrl->sbox, sinit, sizeof(sinit), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemsetAsync(&s->B.ctrl->sbox_pending, 0, sizeof(unsigned), s->stream));
    CU(cudaStreamSynchronize(s->stream));  // init[] is on the stack
    k_bbox_reduce<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B);
    CU(cudaGetLastError());
    return 0;
}

int export_state(dem_b200_system* s) {
    if (s->export_valid)
        return 0;
    k_export_state<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, s->d_pos3, s->d_vel3, s->d_om3);
    CU(cudaGetLastError());
    s->export_valid = true;
    return 0;
}

int run_steps(dem_b200_system* s, int nsteps) {
    if (!s->initialized) {
        s->err = "step before initialize";
        return DEMB200_EINVAL;
    }
    int done = 0;
    // slab mode: the step right after a rebuild runs un-captured (it is followed by the index remap below); the
    // graph is re-captured after every slab rebuild because the local sphere count (grid sizes) changes
    // callers that advance one step per call (co-simulation, the reference's unit tests: SURVEY Q15) get the graph from
    // their third call on
    if (nsteps == 1 && s->one_step_calls < 3)
        s->one_step_calls++;
    // DEMB200_NO_GRAPH=1: every step as direct launches (profiling: ncu lists the kernels of a step one by one, also where it
    // cannot look into the step graph)
    static const bool no_graph = [] {
        const char* e = getenv("DEMB200_NO_GRAPH");
        return e && e[0] == '1';
    }();
    if (!no_graph && !s->recording && ((!s->mgpu && (nsteps >= 2 || (nsteps == 1 && s->one_step_calls >= 3))) || (s->mgpu && !s->mg_remap_pending))) {
        if (!s->graph1) {
            int rc = build_graph(s);
            if (rc)
                return rc;
        }
        for (; done < nsteps; done++) {
            CU(cudaGraphLaunch(s->graph1, s->stream));
            s->time += s->P.dt;
            s->step_no++;
        }
        s->export_valid = false;
        if (nsteps > 0) {
            s->accel_src_valid = true;
            s->accel_export_valid = false;
        }
    }
    for (; done < nsteps; done++) {
        int rc = enqueue_step(s, nullptr);
        if (rc)
            return rc;
        if (s->mg_remap_pending) {
            // that step sorted the freshly assembled slab: turn pre-sort indices into storage slots for the halo lists
            const unsigned N = s->P.N;
            const unsigned m = std::max(std::max(s->mg_ns[0], s->mg_ns[1]), std::max(s->mg_ng[0], s->mg_ng[1]));
            k_mgpu_invert_perm<<<(N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B);
            if (m)
                k_mgpu_remap<<<(m + 255) / 256, 256, 0, s->stream>>>(s->B, s->mg_n_own, s->mg_ns[0], s->mg_ns[1], s->mg_ng[0], s->mg_ng[1]);
            k_mgpu_bnd_list<<<(N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B);
            s->mg_remap_pending = false;
            if (s->p2p) {  // the first step after a slab rebuild: its result feeds the neighbours' next step
                enqueue_halo_pack(s, s->stream);
                enqueue_vote(s, s->stream);
            }
            CU(cudaGetLastError());
        }
    }
    return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

int dem_b200_create(const dem_b200_config* cfg, dem_b200_system** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return DEMB200_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return DEMB200_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        g_create_error = "bad device ordinal";
        return DEMB200_EINVAL;
    }
    dem_b200_system* s = new dem_b200_system();
    s->cfg = *cfg;
    e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaMallocHost((void**)&s->h_pin, 16 * sizeof(double));
    if (e != cudaSuccess) {
        g_create_error = std::string("CUDA init: ") + cudaGetErrorString(e);
        delete s;
        return DEMB200_ECUDA;
    }
    s->P.nW = 0;
    refresh_params(s);
    *out = s;
    return 0;
}

void dem_b200_destroy(dem_b200_system* s) {
    if (!s)
        return;
    cudaSetDevice(s->cfg.device);
    if (s->stream)
        cudaStreamSynchronize(s->stream);
    drop_graph(s);
    for (void* p : s->allocs)
        cudaFree(p);
    if (s->h_pin)
        cudaFreeHost(s->h_pin);
    if (s->h_mesh_pin)
        cudaFreeHost(s->h_mesh_pin);
    for (int r = 0; r < kMaxRanks; r++)
        if (s->p2p_mapped[r])
            cudaIpcCloseMemHandle(s->p2p_mapped[r]);
    if (s->p2p_region)
        cudaFree(s->p2p_region);
    if (s->h_vote)
        cudaFreeHost(s->h_vote);
    if (s->stream && s->own_stream)
        cudaStreamDestroy(s->stream);
    if (s->cap_stream)
        cudaStreamDestroy(s->cap_stream);
    if (s->cap_stream2)
        cudaStreamDestroy(s->cap_stream2);
    if (s->side_stream) {
        cudaStreamDestroy(s->side_stream);
        cudaEventDestroy(s->ev_fork);
        cudaEventDestroy(s->ev_join);
    }
    delete s;
}

const char* dem_b200_last_error(const dem_b200_system* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int dem_b200_set_config(dem_b200_system* s, const dem_b200_config* cfg) {
    if (!s || !cfg)
        return DEMB200_EINVAL;
    if (s->initialized) {
        const dem_b200_config& o = s->cfg;
        if (o.history_slots != cfg->history_slots || o.neighbor_slots != cfg->neighbor_slots || o.device != cfg->device ||
            (o.integrator != cfg->integrator && (o.integrator == DEMB200_CHUNG || cfg->integrator == DEMB200_CHUNG)) ||
            o.tangential_mode != cfg->tangential_mode || o.force_model != cfg->force_model ||
            o.use_mat_props != cfg->use_mat_props) {
            s->err = "set_config: history_slots, neighbor_slots, device, tangential_mode, force_model, use_mat_props and "
                     "(to/from) Chung cannot change after initialize";
            return DEMB200_EINVAL;
        }
    }
    const double old_skin = s->P.skin;
    s->cfg = *cfg;
    refresh_params(s);
    drop_graph(s);
    if (s->initialized && s->P.skin > old_skin) {
        // a larger skin needs longer candidate lists: rebuild at the next step
        CU(cudaSetDevice(s->cfg.device));
        const unsigned one = 1;
        CU(cudaMemcpyAsync(&s->B.ctrl->need_rebuild, &one, sizeof(unsigned), cudaMemcpyHostToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return 0;
}

int dem_b200_set_spheres(dem_b200_system* s, size_t n, const double* pos3, const double* vel3, const double* omega3,
                         const double* radius, const uint8_t* fixed) {
    if (!s || !pos3 || !radius || n == 0 || n >= 0xFFFF0000ull) {
        if (s) s->err = "set_spheres: bad arguments";
        return DEMB200_EINVAL;
    }
    if (s->initialized) {
        s->err = "set_spheres after initialize";
        return DEMB200_EINVAL;
    }
    s->h_pos.assign(pos3, pos3 + 3 * n);
    s->h_rad.assign(radius, radius + n);
    if (vel3) s->h_vel.assign(vel3, vel3 + 3 * n); else s->h_vel.assign(3 * n, 0.0);
    if (omega3) s->h_om.assign(omega3, omega3 + 3 * n); else s->h_om.assign(3 * n, 0.0);
    s->h_ids.clear();
    s->any_fixed = false;
    s->h_fixed.assign(n, 0);
    if (fixed)
        for (size_t i = 0; i < n; i++) {
            s->h_fixed[i] = fixed[i] ? 1 : 0;
            s->any_fixed |= (fixed[i] != 0);
        }
    return 0;
}

static int add_wall(dem_b200_system* s, int type, const double pos[3], const double rot[4], const double hd[3]) {
    if (!s || !pos || !hd)
        return DEMB200_EINVAL;
    if (s->initialized) {
        s->err = "walls must be added before initialize";
        return DEMB200_EINVAL;
    }
    if (s->P.nW >= kMaxWalls) {
        s->err = "too many walls";
        return DEMB200_EINVAL;
    }
    Wall& W = s->W.w[s->P.nW];
    memset(&W, 0, sizeof(W));
    W.type = type;
    W.enabled = 1;
    for (int k = 0; k < 3; k++) {
        W.pos[k] = pos[k];
        W.hdims[k] = hd[k];
    }
    W.rot[0] = 1;
    if (rot)
        for (int k = 0; k < 4; k++)
            W.rot[k] = rot[k];
    s->P.nW++;
    refresh_params(s);
    return s->P.nW - 1;
}

int dem_b200_add_box_wall(dem_b200_system* s, const double pos[3], const double rot[4], const double hdims[3]) {
    return add_wall(s, WALL_BOX, pos, rot, hdims);
}
int dem_b200_add_plane_wall(dem_b200_system* s, const double pos[3], const double normal[3]) {
    if (!normal)
        return DEMB200_EINVAL;
    double l = std::sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
    if (!(l > 0))
        return DEMB200_EINVAL;
    double n[3] = {normal[0] / l, normal[1] / l, normal[2] / l};
    return add_wall(s, WALL_PLANE, pos, nullptr, n);
}
int dem_b200_add_zcylinder_wall(dem_b200_system* s, const double center[3], double radius, int spheres_inside) {
    if (!(radius > 0))
        return DEMB200_EINVAL;
    double hd[3] = {radius, spheres_inside ? 1.0 : -1.0, 0.0};
    return add_wall(s, WALL_ZCYL, center, nullptr, hd);
}

int dem_b200_add_sphere_wall(dem_b200_system* s, const double center[3], double radius, int spheres_outside) {
    if (!(radius > 0))
        return DEMB200_EINVAL;
    double hd[3] = {radius, spheres_outside ? 1.0 : -1.0, 0.0};
    return add_wall(s, WALL_SPHERE, center, nullptr, hd);
}
int dem_b200_add_zcone_wall(dem_b200_system* s, const double tip[3], double slope, double hmin, double hmax, int spheres_above) {
    if (!(slope > 0) || !(hmax > hmin))
        return DEMB200_EINVAL;
    double hd[3] = {slope, hmin, hmax};
    double side[4] = {spheres_above ? 1.0 : -1.0, 0.0, 0.0, 0.0};
    return add_wall(s, WALL_ZCONE, tip, side, hd);
}

static int upload_walls(dem_b200_system* s) {
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemcpyAsync(s->B.walls, &s->W, sizeof(WallSet), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));  // s->W may change again right after this call
    return 0;
}

int dem_b200_set_wall_state(dem_b200_system* s, int w, const double pos[3], const double vel[3]) {
    if (!s || w < 0 || w >= s->P.nW)
        return DEMB200_EINVAL;
    double moved = 0;
    for (int k = 0; k < 3; k++) {
        if (pos) {
            moved += (pos[k] - s->W.w[w].pos[k]) * (pos[k] - s->W.w[w].pos[k]);
            s->W.w[w].pos[k] = pos[k];
        }
        if (vel) s->W.w[w].vel[k] = vel[k];
    }
    refresh_params(s);
    if (!s->initialized)
        return 0;
    // walls live in device memory: the captured graph stays valid.  A moving wall uses up Verlet skin like a moving
    // sphere (the per-sphere wall-candidate masks were computed with a skin/2 reach).
    int rc = upload_walls(s);
    if (rc)
        return rc;
    if (moved > 0) {
        k_wall_moved<<<1, 1, 0, s->stream>>>(s->B, std::sqrt(moved));
        CU(cudaGetLastError());
    }
    return 0;
}
int dem_b200_set_wall_velocity(dem_b200_system* s, int w, const double pos[3], const double vel[3]) {
    return dem_b200_set_wall_state(s, w, pos, vel);
}
int dem_b200_enable_wall(dem_b200_system* s, int w, int enabled) {
    if (!s || w < 0 || w >= s->P.nW)
        return DEMB200_EINVAL;
    s->W.w[w].enabled = enabled ? 1 : 0;
    return s->initialized ? upload_walls(s) : 0;
}
int dem_b200_track_wall_forces(dem_b200_system* s, int enable) {
    if (!s)
        return DEMB200_EINVAL;
    s->track_wall_forces = enable != 0;
    refresh_params(s);
    drop_graph(s);
    return 0;
}
int dem_b200_wall_force(dem_b200_system* s, int w, double force[3]) {
    if (!s || !s->initialized || w < 0 || w >= s->P.nW || !force)
        return DEMB200_EINVAL;
    if (!s->track_wall_forces) {
        s->err = "wall_force: call dem_b200_track_wall_forces(s, 1) first";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemcpyAsync(s->h_pin, s->B.ctrl->wall_force[w], 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < 3; k++)
        force[k] = s->h_pin[k];
    return 0;
}
int dem_b200_set_contact_class(dem_b200_system* s, int cls, const dem_b200_contact_class* c) {
    if (!s || cls < 0 || cls > 2)
        return DEMB200_EINVAL;
    s->cls_override[cls] = (c != nullptr);
    if (c)
        s->cls[cls] = *c;
    refresh_params(s);
    drop_graph(s);
    return 0;
}
int dem_b200_num_walls(const dem_b200_system* s) { return s ? s->P.nW : 0; }

// ---- triangle meshes (ChSystemDemMesh::AddMesh / ApplyMeshMotion / CollectMeshContactForces) --------------------
int dem_b200_add_mesh(dem_b200_system* s, size_t ntri, const double* verts9, double mass) {
    if (!s || !verts9 || ntri == 0)
        return DEMB200_EINVAL;
    if (s->initialized) {
        s->err = "meshes must be added before initialize";
        return DEMB200_EINVAL;
    }
    if (s->M.n >= kMaxMeshes || s->h_tri_mesh.size() + ntri >= 0x7FFF0000ull) {
        s->err = "too many meshes / triangles";
        return DEMB200_EINVAL;
    }
    const int m = s->M.n++;
    MeshBody& B = s->M.m[m];
    memset(&B, 0, sizeof(B));
    B.rot[0] = 1.0;
    B.mass = mass > 0 ? mass : s->cfg.mesh_mass;
    B.tri_begin = (unsigned)s->h_tri_mesh.size();
    B.tri_end = B.tri_begin + (unsigned)ntri;
    s->M.enabled = 1;
    s->h_tri.insert(s->h_tri.end(), verts9, verts9 + 9 * ntri);
    s->h_tri_mesh.insert(s->h_tri_mesh.end(), ntri, (uint32_t)m);
    double r2 = 0;
    for (size_t i = 0; i < 3 * ntri; i++)
        r2 = std::max(r2, verts9[3 * i] * verts9[3 * i] + verts9[3 * i + 1] * verts9[3 * i + 1] + verts9[3 * i + 2] * verts9[3 * i + 2]);
    s->mesh_radius[m] = std::sqrt(r2);
    refresh_params(s);
    return m;
}

int dem_b200_num_meshes(const dem_b200_system* s) { return s ? s->M.n : 0; }
size_t dem_b200_num_triangles(const dem_b200_system* s) { return s ? s->h_tri_mesh.size() : 0; }

int dem_b200_set_mesh_motion(dem_b200_system* s, int m, const double pos[3], const double rot[4], const double lin_vel[3],
                             const double ang_vel[3]) {
    if (!s || m < 0 || m >= s->M.n)
        return DEMB200_EINVAL;
    MeshBody& B = s->M.m[m];
    // how far can a vertex have moved?  |dx| + (rotation angle between the two frames) * (largest vertex radius)
    double moved = 0;
    if (pos) {
        double d2 = 0;
        for (int k = 0; k < 3; k++) {
            d2 += (pos[k] - B.pos[k]) * (pos[k] - B.pos[k]);
            B.pos[k] = pos[k];
        }
        moved += std::sqrt(d2);
    }
    if (rot) {
        double l = std::sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3]);
        if (!(l > 0))
            return DEMB200_EINVAL;
        double dq = 0;
        for (int k = 0; k < 4; k++)
            dq += rot[k] / l * B.rot[k];
        dq = std::min(1.0, std::fabs(dq));
        const double ang = 2.0 * std::acos(dq);
        // acos loses accuracy next to 1: bound the angle from the chord |q - q'| as well
        double c2 = 0;
        const double sg = (rot[0] * B.rot[0] + rot[1] * B.rot[1] + rot[2] * B.rot[2] + rot[3] * B.rot[3]) < 0 ? -1.0 : 1.0;
        for (int k = 0; k < 4; k++)
            c2 += (sg * rot[k] / l - B.rot[k]) * (sg * rot[k] / l - B.rot[k]);
        moved += std::max(ang, 2.2 * std::sqrt(c2)) * s->mesh_radius[m];
        for (int k = 0; k < 4; k++)
            B.rot[k] = rot[k];  // as given (the oracle / Multicore use the body quaternion unnormalised too)
    }
    for (int k = 0; k < 3; k++) {
        if (lin_vel) B.vel[k] = lin_vel[k];
        if (ang_vel) B.omg[k] = ang_vel[k];
    }
    if (!s->initialized)
        return 0;
    CU(cudaSetDevice(s->cfg.device));
    // asynchronous on the engine's stream; a staging slot is re-used kMeshRing calls later, so the host only waits when
    // the ring wraps
    const unsigned slot = s->mesh_ring_pos++ % kMeshRing;
    if (slot == 0 && s->mesh_ring_pos > 1)
        CU(cudaStreamSynchronize(s->stream));
    MeshBody* stage = s->h_mesh_pin + (size_t)slot * kMaxMeshes;
    memcpy(stage, s->M.m, sizeof(s->M.m));
    CU(cudaMemcpyAsync(s->B.meshes->m, stage, sizeof(s->M.m), cudaMemcpyHostToDevice, s->stream));
    if (pos || rot) {
        const unsigned nt = B.tri_end - B.tri_begin;
        k_mesh_begin<<<1, 32, 0, s->stream>>>(s->B, m);
        k_mesh_transform<<<(nt + 255) / 256, 256, 0, s->stream>>>(s->B, m);
        if (moved > 0)
            k_mesh_moved<<<1, 1, 0, s->stream>>>(s->B, moved);
        CU(cudaGetLastError());
    }
    return 0;
}

int dem_b200_enable_mesh_collision(dem_b200_system* s, int enabled) {
    if (!s)
        return DEMB200_EINVAL;
    s->M.enabled = enabled ? 1 : 0;
    if (!s->initialized || !s->P.nT)
        return 0;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemcpyAsync(&s->B.meshes->enabled, &s->M.enabled, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    const unsigned one = 1;  // the candidate lists hold no triangles while collision is off
    CU(cudaMemcpyAsync(&s->B.ctrl->need_rebuild, &one, sizeof(unsigned), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int dem_b200_mesh_wrench(dem_b200_system* s, int m, double force[3], double torque[3]) {
    if (!s || !s->initialized || m < 0 || m >= s->M.n)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemcpyAsync(s->h_pin, s->B.meshes->wrench[m], 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < 3; k++) {
        if (force) force[k] = s->h_pin[k];
        if (torque) torque[k] = s->h_pin[3 + k];
    }
    return 0;
}

int dem_b200_initialize(dem_b200_system* s) {
    if (!s)
        return DEMB200_EINVAL;
    if (s->initialized) {
        s->err = "initialize called twice";
        return DEMB200_EINVAL;
    }
    const size_t n = s->h_rad.size();
    if (n == 0) {
        s->err = "initialize without spheres";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    Params& P = s->P;
    Buffers& B = s->B;
    P.N = (unsigned)n;
    if (s->mgpu && s->mg_cap < n) {
        s->err = "mgpu: capacity smaller than the number of spheres";
        return DEMB200_EINVAL;
    }
    P.Np = (unsigned)(((s->mgpu ? s->mg_cap : n) + 31) / 32 * 32);
    P.rmax = *std::max_element(s->h_rad.begin(), s->h_rad.end());
    if (s->mgpu && s->mg_rmax > P.rmax)
        P.rmax = s->mg_rmax;
    if (!s->h_ids.empty() && s->h_ids.size() != n) {
        s->err = "set_sphere_ids: size differs from set_spheres";
        return DEMB200_EINVAL;
    }
    if (!s->h_ids.empty() && !s->mgpu) {
        // without slab mode the ids index user-order buffers of n entries (get_state, recording ...): a permutation of 0 .. n-1
        std::vector<uint8_t> seen(n, 0);
        for (uint32_t id : s->h_ids) {
            if (id >= n || seen[id]) {
                s->err = "set_sphere_ids: outside slab mode the ids must be a permutation of 0 .. n-1";
                return DEMB200_EINVAL;
            }
            seen[id] = 1;
        }
    }
    refresh_params(s);
    if (P.bins[0] < 1 || P.bins[1] < 1 || P.bins[2] < 1) {
        s->err = "bad bins_per_axis";
        return DEMB200_EINVAL;
    }
    if (!(P.rmax > 0) || !(P.skin >= 0)) {
        s->err = "bad radius / verlet_skin";
        return DEMB200_EINVAL;
    }
    // Tiled search grid + tile force kernel (shared-memory staging of the neighbour bins): on by default where the tile kernel
    // applies (Hertz fast paths, no meshes); DEMB200_TILE=0 in the environment keeps the plain grid and k_force_integrate.
    {
        const char* e = getenv("DEMB200_TILE");
        const bool want = e ? (e[0] != '0') : (DEMB200_TILE_DEFAULT != 0);
        P.tiled = (want && !P.nT && fast_path(s)) ? 1 : 0;
    }
    // capacity of the search grid: twice the cells of the initial bounding box (spheres + walls); if the bed ever
    // spreads beyond that, k_step_begin coarsens the cells instead of overflowing
    {
        double mn[3], mx[3];
        for (int k = 0; k < 3; k++) {  // the search grid covers the spheres and the meshes, not the walls
            mn[k] = INFINITY;
            mx[k] = -INFINITY;
        }
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
                mn[k] = std::min(mn[k], s->h_pos[3 * i + k] - s->h_rad[i]);
                mx[k] = std::max(mx[k], s->h_pos[3 * i + k] + s->h_rad[i]);
            }
        for (int m = 0; m < s->M.n; m++)  // a mesh can be turned about its origin: reserve for its bounding sphere
            for (int k = 0; k < 3; k++) {
                mn[k] = std::min(mn[k], s->M.m[m].pos[k] - s->mesh_radius[m]);
                mx[k] = std::max(mx[k], s->M.m[m].pos[k] + s->mesh_radius[m]);
            }
        const double e = 2.0 * P.rmax + P.skin;
        double cells = 1.0;
        for (int k = 0; k < 3; k++)
            cells *= std::max(1.0, std::floor((mx[k] - mn[k]) / e) + 1.0);
        if (P.tiled) {  // border tiles are only partly inside the box: count whole tiles
            cells = 1.0;
            for (int k = 0; k < 3; k++)
                cells *= (std::floor((std::max(1.0, std::floor((mx[k] - mn[k]) / e) + 1.0) + kTile - 1) / kTile) + 1.0) * kTile;
        }
        cells = std::min(std::max(2.0 * cells, 4096.0), std::max(4096.0, 16.0 * (double)(s->mgpu ? s->mg_cap : n)));
        P.cell_cap = (unsigned)cells;
    }
    s->ntiles = (P.cell_cap + kScanTile - 1) / kScanTile;
    const size_t K = (size_t)P.K, Np = P.Np;
    const bool hist = (P.tang_mode == DEMB200_TANG_MULTISTEP);
    s->use_hrel = hist && P.use_mat_props && (P.force_model == DEMB200_HOOKE || P.force_model == DEMB200_FLORES);

    int rc = 0;
    const size_t HS = (size_t)P.Kn + (size_t)P.nW;  // history slots per sphere: one per candidate + one per wall
    rc |= dev_alloc(s, &B.ctrl, 1);
    rc |= dev_alloc(s, &B.walls, 1);
    for (int b = 0; b < 2; b++) {
        rc |= dev_alloc(s, &B.pos[b], Np);
        rc |= dev_alloc(s, &B.vel[b], Np);
        if (P.integrator == DEMB200_CHUNG)
            rc |= dev_alloc(s, &B.acc[b], 6 * Np);
    }
    if (hist) {
        rc |= dev_alloc(s, &B.hist, HS * Np);
        rc |= dev_alloc(s, &B.stage, K * Np);
        rc |= dev_alloc(s, &B.stage_cnt, Np);
        if (s->use_hrel) {
            rc |= dev_alloc(s, &B.hrel, HS * Np);
            rc |= dev_alloc(s, &B.stage_rel, K * Np);
        }
    }
    rc |= dev_alloc(s, &B.cell, Np); rc |= dev_alloc(s, &B.rank, Np); rc |= dev_alloc(s, &B.perm, Np);
    rc |= dev_alloc(s, &B.cell_count, (size_t)P.cell_cap + 8); rc |= dev_alloc(s, &B.cell_start, (size_t)P.cell_cap + 8);
    rc |= dev_alloc(s, &B.block_sums, (size_t)s->ntiles + 8);
    rc |= dev_alloc(s, &B.nl, (size_t)P.Kn * Np); rc |= dev_alloc(s, &B.ncnt, Np);
    if (P.tiled)
        rc |= dev_alloc(s, &B.nl16, (size_t)P.Kn * Np);
    rc |= dev_alloc(s, &s->d_pos3, 3 * Np); rc |= dev_alloc(s, &s->d_vel3, 3 * Np); rc |= dev_alloc(s, &s->d_om3, 3 * Np);
    rc |= dev_alloc(s, &s->d_red, 4);
    if (P.nT) {
        // capacity of the (cell, triangle) list: every triangle reaches at most the cells of the cube around its longest
        // edge inflated by the reach of a sphere
        const double e = 2.0 * P.rmax + P.skin, reach = P.rmax + 0.5 * P.skin_tri;
        double tot = 0;
        for (size_t t = 0; t < P.nT; t++) {
            const double* v = &s->h_tri[9 * t];
            double L = 0;
            for (int a = 0; a < 3; a++) {
                const int b = (a + 1) % 3;
                double d2 = 0;
                for (int k = 0; k < 3; k++)
                    d2 += (v[3 * a + k] - v[3 * b + k]) * (v[3 * a + k] - v[3 * b + k]);
                L = std::max(L, std::sqrt(d2));
            }
            const double c = std::floor((L + 2 * reach) / e) + 2.0;
            tot += std::min(c * c * c, 20.0 * c * c);  // the plane cull keeps a slab of ~3-4 cells (+ grid boundary cells)
        }
        P.tri_cap = (unsigned)std::min(std::max(2.0 * tot, 1024.0), 1.0e9);
        rc |= dev_alloc(s, &B.meshes, 1);
        rc |= dev_alloc(s, &B.tri_loc, 9 * (size_t)P.nT); rc |= dev_alloc(s, &B.tri_w, 9 * (size_t)P.nT);
        rc |= dev_alloc(s, &B.tri_mesh, (size_t)P.nT);
        rc |= dev_alloc(s, &B.tcell_count, (size_t)P.cell_cap + 8); rc |= dev_alloc(s, &B.tcell_start, (size_t)P.cell_cap + 8);
        rc |= dev_alloc(s, &B.tblock_sums, (size_t)s->ntiles + 8);
        rc |= dev_alloc(s, &B.tcell_tri, (size_t)P.tri_cap);
    }
    if (s->mgpu) {
        rc |= dev_alloc(s, &B.slab, 1);
        rc |= dev_alloc(s, &B.inv_perm, Np);
        rc |= dev_alloc(s, &B.bnd_list, Np);
        rc |= dev_alloc(s, &s->d_count, 4);
        for (int d = 0; d < 2; d++) {
            rc |= dev_alloc(s, &B.send_pre[d], Np); rc |= dev_alloc(s, &B.send_slot[d], Np); rc |= dev_alloc(s, &B.ghost_slot[d], Np);
        }
        if (hist) {
            rc |= dev_alloc(s, &B.stage_init, K * Np);
            rc |= dev_alloc(s, &B.stage_cnt_init, Np);
            if (s->use_hrel)
                rc |= dev_alloc(s, &B.stage_rel_init, K * Np);
        }
    }
    if (rc)
        return DEMB200_ECUDA;
    if (s->mgpu) {
        CU(cudaMemset(B.slab, 0, sizeof(SlabDev)));
        if (B.stage_cnt_init)
            CU(cudaMemset(B.stage_cnt_init, 0, Np * sizeof(uint32_t)));
    }

    // history supplied before initialize (checkpoint restart): every sphere taking part in a contact gets a record
    // (keyed by the partner's shape id); the first rebuild moves them into the candidate slots
    std::vector<std::vector<dem_b200_system::HRow>> rows;
    if (hist && !s->h_hist.empty()) {
        rows.resize(n);
        for (auto& r : s->h_hist) {
            if (r.owner < P.shape_base || r.owner - P.shape_base >= n) {
                s->err = "add_history: owner is not a sphere shape";
                return DEMB200_EINVAL;
            }
            const size_t so = r.owner - P.shape_base;
            rows[so].push_back(r);  // key = other
            if (r.other >= P.shape_base) {
                if (r.other - P.shape_base >= n) {
                    s->err = "add_history: bad partner shape id";
                    return DEMB200_EINVAL;
                }
                dem_b200_system::HRow m = r;
                m.other = r.owner;  // the partner's copy is keyed by the owner's shape id
                rows[r.other - P.shape_base].push_back(m);
            }
        }
        for (auto& v : rows)
            if (v.size() > K) {
                s->err = "add_history: too many rows for one sphere";
                return DEMB200_EHISTORY;
            }
        if (!B.stage_init) {
            rc |= dev_alloc(s, &B.stage_init, K * Np);
            rc |= dev_alloc(s, &B.stage_cnt_init, Np);
            if (s->use_hrel)
                rc |= dev_alloc(s, &B.stage_rel_init, K * Np);
        }
        if (rc)
            return DEMB200_ECUDA;
    }

    // upload (storage order = user order initially; the first step sorts by search cell)
    {
        std::vector<double4> hp(Np, make_double4(0, 0, 0, 1));
        std::vector<VelRec> hv(Np);
        memset(hv.data(), 0, Np * sizeof(VelRec));
        for (size_t i = 0; i < n; i++) {
            hp[i] = make_double4(s->h_pos[3 * i], s->h_pos[3 * i + 1], s->h_pos[3 * i + 2], s->h_rad[i]);
            for (int k = 0; k < 3; k++) {
                hv[i].v[k] = s->h_vel[3 * i + k];
                hv[i].w[k] = s->h_om[3 * i + k];
            }
            hv[i].sid = s->h_ids.empty() ? (uint32_t)i : s->h_ids[i];
            hv[i].meta = s->h_fixed[i] ? 1u : 0u;
            hv[i].amask = 0ull;
        }
        for (int b = 0; b < 2; b++) {
            CU(cudaMemcpy(B.pos[b], hp.data(), Np * sizeof(double4), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(B.vel[b], hv.data(), Np * sizeof(VelRec), cudaMemcpyHostToDevice));
            if (B.acc[b])
                CU(cudaMemset(B.acc[b], 0, 6 * Np * sizeof(double)));
        }
        if (B.hist) {
            CU(cudaMemset(B.hist, 0, HS * Np * sizeof(double4)));
            CU(cudaMemset(B.stage_cnt, 0, Np * sizeof(uint32_t)));
        }
        if (B.hrel)
            CU(cudaMemset(B.hrel, 0, HS * Np * sizeof(double)));
    }
    if (!rows.empty()) {
        std::vector<double4> hh(K * Np, make_double4(0, 0, 0, 0));
        std::vector<double> hr(s->use_hrel ? K * Np : 0, 0.0);
        std::vector<uint32_t> hcn(Np, 0);
        for (size_t i = 0; i < n; i++) {
            hcn[i] = (uint32_t)rows[i].size();
            for (size_t k = 0; k < rows[i].size(); k++) {
                const auto& r = rows[i][k];
                const unsigned steps = (unsigned)std::llround(r.dur / P.dt);
                const unsigned long long bits = ((unsigned long long)steps << 32) | r.other;
                double w;
                memcpy(&w, &bits, 8);
                hh[k * Np + i] = make_double4(r.d[0], r.d[1], r.d[2], w);
                if (s->use_hrel)
                    hr[k * Np + i] = r.rel;
            }
        }
        CU(cudaMemcpy(B.stage_init, hh.data(), K * Np * sizeof(double4), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(B.stage_cnt_init, hcn.data(), Np * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (s->use_hrel)
            CU(cudaMemcpy(B.stage_rel_init, hr.data(), K * Np * sizeof(double), cudaMemcpyHostToDevice));
    }
    CU(cudaMemset(B.cell_count, 0, ((size_t)P.cell_cap + 8) * sizeof(uint32_t)));
    CU(cudaMemset(B.ncnt, 0, Np * sizeof(uint32_t)));
    if (P.nT) {
        CU(cudaMemset(B.tcell_count, 0, ((size_t)P.cell_cap + 8) * sizeof(uint32_t)));
        CU(cudaMemcpy(B.tri_loc, s->h_tri.data(), 9 * (size_t)P.nT * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(B.tri_mesh, s->h_tri_mesh.data(), (size_t)P.nT * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(B.meshes, &s->M, sizeof(MeshSet), cudaMemcpyHostToDevice));
        CU(cudaMallocHost((void**)&s->h_mesh_pin, kMeshRing * sizeof(s->M.m)));
        for (int m = 0; m < s->M.n; m++) {
            const unsigned nt = s->M.m[m].tri_end - s->M.m[m].tri_begin;
            k_mesh_begin<<<1, 32, 0, s->stream>>>(B, m);
            k_mesh_transform<<<(nt + 255) / 256, 256, 0, s->stream>>>(B, m);
        }
        CU(cudaGetLastError());
    }
    {
        Ctrl c;
        memset(&c, 0, sizeof(c));
        c.cur = 0;
        c.need_rebuild = 1;
        c.init_stage = rows.empty() ? 0u : 1u;
        CU(cudaMemcpy(B.ctrl, &c, sizeof(Ctrl), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(B.walls, &s->W, sizeof(WallSet), cudaMemcpyHostToDevice));
    }
    s->h_pos.clear(); s->h_pos.shrink_to_fit();
    s->h_vel.clear(); s->h_vel.shrink_to_fit();
    s->h_om.clear(); s->h_om.shrink_to_fit();
    s->initialized = true;
    s->mg_n_own = s->mg_n_local = (unsigned)n;
    rc = recompute_bbox(s);
    if (rc)
        return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int dem_b200_step(dem_b200_system* s, int nsteps) {
    if (!s || nsteps < 0)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    return run_steps(s, nsteps);
}

int dem_b200_sync(dem_b200_system* s) {
    if (!s)
        return DEMB200_EINVAL;
    if (!s->initialized) {
        s->err = "sync before initialize";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    return check_device_error(s);
}

int dem_b200_step_timed(dem_b200_system* s, int nsteps, float* ms) {
    if (!s || !ms)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, s->stream));
    int rc = run_steps(s, nsteps);
    CU(cudaEventRecord(b, s->stream));
    CU(cudaEventSynchronize(b));
    CU(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (rc)
        return rc;
    return check_device_error(s);
}

int dem_b200_step_profile(dem_b200_system* s, int nsteps, float* ms_per_kernel, int* n_out) {
    if (!s || !ms_per_kernel || !n_out || !s->initialized)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    cudaEvent_t ev[kNumKernels + 1];
    for (auto& e : ev)
        CU(cudaEventCreate(&e));
    for (int k = 0; k < kNumKernels; k++)
        ms_per_kernel[k] = 0.f;
    for (int i = 0; i < nsteps; i++) {
        int rc = enqueue_step(s, ev);
        if (rc)
            return rc;
        CU(cudaEventSynchronize(ev[kNumKernels]));
        for (int k = 0; k < kNumKernels; k++) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
            ms_per_kernel[k] += t;
        }
    }
    drop_graph(s);
    for (auto& e : ev)
        cudaEventDestroy(e);
    *n_out = kNumKernels;
    return check_device_error(s);
}

const char* dem_b200_kernel_name(int k) { return (k >= 0 && k < kNumKernels) ? kKernelNames[k] : ""; }

size_t dem_b200_num_spheres(const dem_b200_system* s) { return s ? (s->initialized ? s->P.N : s->h_rad.size()) : 0; }
double dem_b200_time(const dem_b200_system* s) { return s ? s->time : 0.0; }

int dem_b200_get_state(dem_b200_system* s, double* pos3, double* vel3, double* omega3) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    int rc = export_state(s);
    if (rc)
        return rc;
    const size_t bytes = 3 * (size_t)s->P.N * sizeof(double);
    if (pos3) CU(cudaMemcpyAsync(pos3, s->d_pos3, bytes, cudaMemcpyDeviceToHost, s->stream));
    if (vel3) CU(cudaMemcpyAsync(vel3, s->d_vel3, bytes, cudaMemcpyDeviceToHost, s->stream));
    if (omega3) CU(cudaMemcpyAsync(omega3, s->d_om3, bytes, cudaMemcpyDeviceToHost, s->stream));
    return check_device_error(s);
}

int dem_b200_get_sphere(dem_b200_system* s, size_t i, double pos[3], double vel[3], double omega[3]) {
    if (!s || !s->initialized || i >= s->P.N)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    int rc = export_state(s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(s->h_pin + 1, s->d_pos3 + 3 * i, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(s->h_pin + 4, s->d_vel3 + 3 * i, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(s->h_pin + 7, s->d_om3 + 3 * i, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    rc = check_device_error(s);  // synchronises
    for (int k = 0; k < 3; k++) {
        if (pos) pos[k] = s->h_pin[1 + k];
        if (vel) vel[k] = s->h_pin[4 + k];
        if (omega) omega[k] = s->h_pin[7 + k];
    }
    return rc;
}

int dem_b200_set_state(dem_b200_system* s, const double* pos3, const double* vel3, const double* omega3) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    const size_t bytes = 3 * (size_t)s->P.N * sizeof(double);
    if (pos3) CU(cudaMemcpyAsync(s->d_pos3, pos3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (vel3) CU(cudaMemcpyAsync(s->d_vel3, vel3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (omega3) CU(cudaMemcpyAsync(s->d_om3, omega3, bytes, cudaMemcpyHostToDevice, s->stream));
    k_import_state<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, pos3 ? s->d_pos3 : nullptr,
                                                                vel3 ? s->d_vel3 : nullptr, omega3 ? s->d_om3 : nullptr);
    CU(cudaGetLastError());
    s->export_valid = false;
    s->accel_src_valid = s->accel_export_valid = false;
    if (pos3)
        return recompute_bbox(s);
    return 0;
}

int dem_b200_request_rebuild(dem_b200_system* s) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "slab engines rebuild through the slab protocol (dem_b200_p2p_rebuild / mgpu_finish_rebuild)";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    const unsigned one = 1;
    CU(cudaMemcpyAsync(&s->B.ctrl->need_rebuild, &one, sizeof(unsigned), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));  // `one` is on the stack
    return 0;
}

namespace {
int export_accel(dem_b200_system* s) {
    if (s->mgpu)
        return DEMB200_EINVAL;  // slab mode: the owned set changes at every slab rebuild; not offered there
    const size_t N = s->P.N;
    if (!s->d_acc3) {
        int rc = dev_alloc(s, &s->d_acc3, 3 * N);
        if (rc)
            return rc;
    }
    if (s->accel_export_valid)
        return 0;
    if (!s->accel_src_valid) {  // no step since Initialize / set_state: zero, as the reference's freshly reset sphere_acc
        CU(cudaMemsetAsync(s->d_acc3, 0, 3 * N * sizeof(double), s->stream));
    } else {
        k_export_accel<<<(unsigned)((N + 255) / 256), 256, 0, s->stream>>>(s->P, s->B, s->d_acc3);
        CU(cudaGetLastError());
    }
    s->accel_export_valid = true;
    return 0;
}
}  // namespace

int dem_b200_get_accel(dem_b200_system* s, double* acc3) {
    if (!s || !s->initialized || !acc3)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    int rc = export_accel(s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(acc3, s->d_acc3, 3 * (size_t)s->P.N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    return check_device_error(s);
}

int dem_b200_get_sphere_accel(dem_b200_system* s, size_t i, double acc[3]) {
    if (!s || !s->initialized || !acc || i >= s->P.N)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    int rc = export_accel(s);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(s->h_pin + 1, s->d_acc3 + 3 * i, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    rc = check_device_error(s);  // synchronises
    memcpy(acc, s->h_pin + 1, 3 * sizeof(double));
    return rc;
}

int dem_b200_advance_host(dem_b200_system* s, size_t n, const double* pos3_in, const double* vel3_in,
                          const double* omega3_in, int nsteps, double* pos3_out, double* vel3_out,
                          double* omega3_out) {
    if (!s || !s->initialized || n != s->P.N)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    int rc = 0;
    if (pos3_in || vel3_in || omega3_in)
        rc = dem_b200_set_state(s, pos3_in, vel3_in, omega3_in);
    if (rc)
        return rc;
    rc = run_steps(s, nsteps);
    if (rc)
        return rc;
    return dem_b200_get_state(s, pos3_out, vel3_out, omega3_out);
}

int dem_b200_reduce(dem_b200_system* s, int which, double arg, double* out) {
    if (!s || !s->initialized || !out)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    const unsigned N = s->P.N;
    if (which < 0 || which > 7)
        return DEMB200_EINVAL;
    if (which == DEMB200_RED_NUM_CONTACTS && s->P.tang_mode != DEMB200_TANG_MULTISTEP) {
        s->err = "NUM_CONTACTS needs MultiStep history (or use recording)";
        return DEMB200_EINVAL;
    }
    unsigned long long init[2] = {0ull, enc_ord_h(-INFINITY)};
    memcpy(s->h_pin + 12, init, 16);
    CU(cudaMemcpyAsync(s->d_red, s->h_pin + 12, 16, cudaMemcpyHostToDevice, s->stream));
    k_reduce<<<(N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, which, arg, s->d_red,
                                                      (unsigned long long*)(s->d_red + 1));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->h_pin, s->d_red, 16, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (which == 2 || which >= 4) {
        *out = s->h_pin[0];
    } else {
        unsigned long long u;
        memcpy(&u, s->h_pin + 1, 8);
        u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
        double d;
        memcpy(&d, &u, 8);
        *out = (which == 1) ? -d : d;
    }
    return 0;
}

int dem_b200_enable_recording(dem_b200_system* s, int enable, size_t max_pairs) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    drop_graph(s);
    s->user_recording = enable != 0;
    s->recording = enable != 0 || s->contact_info;
    if (!enable)
        return 0;
    const size_t n = s->P.N;
    if (!s->B.recF) {
        int rc = 0;
        rc |= dev_alloc(s, &s->B.recF, 3 * n); rc |= dev_alloc(s, &s->B.recT, 3 * n);
        rc |= dev_alloc(s, &s->B.gmin, 3 * n); rc |= dev_alloc(s, &s->B.gmax, 3 * n);
        if (rc)
            return DEMB200_ECUDA;
    }
    if (max_pairs > s->max_pairs) {
        int rc = dev_alloc(s, &s->B.pairs, max_pairs);
        if (rc)
            return DEMB200_ECUDA;
        s->max_pairs = max_pairs;
    }
    s->B.pair_cap = s->max_pairs;
    return 0;
}

int dem_b200_get_forces(dem_b200_system* s, double* force3, double* torque3) {
    if (!s || !s->initialized || !s->B.recF)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    const size_t bytes = 3 * (size_t)s->P.N * sizeof(double);
    if (force3) CU(cudaMemcpyAsync(force3, s->B.recF, bytes, cudaMemcpyDeviceToHost, s->stream));
    if (torque3) CU(cudaMemcpyAsync(torque3, s->B.recT, bytes, cudaMemcpyDeviceToHost, s->stream));
    return check_device_error(s);
}

int dem_b200_get_pairs(dem_b200_system* s, uint64_t* pairs, size_t capacity, size_t* n) {
    if (!s || !s->initialized || !s->B.pairs || !n)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemcpyAsync(s->h_pin, &s->B.ctrl->pair_count, 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    unsigned long long c;
    memcpy(&c, s->h_pin, 8);
    *n = (size_t)c;
    if (c > s->max_pairs) {
        s->err = "pair recording buffer overflow";
        return DEMB200_ECAPACITY;
    }
    if (pairs) {
        if (capacity < c)
            return DEMB200_ECAPACITY;
        CU(cudaMemcpy(pairs, s->B.pairs, c * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int dem_b200_get_bins(dem_b200_system* s, int32_t* gmin3, int32_t* gmax3) {
    if (!s || !s->initialized || !s->B.gmin)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "not available on a slab engine (ids are global, buffers local): use dem_b200_export_owned / dem_b200_import_owned";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    const size_t bytes = 3 * (size_t)s->P.N * sizeof(int32_t);
    if (gmin3) CU(cudaMemcpy(gmin3, s->B.gmin, bytes, cudaMemcpyDeviceToHost));
    if (gmax3) CU(cudaMemcpy(gmax3, s->B.gmax, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int dem_b200_get_grid(dem_b200_system* s, double origin[3], double bin_size[3], double inv_bin_size[3]) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    Ctrl c;
    int rc = read_ctrl(s, &c);
    if (rc)
        return rc;
    for (int k = 0; k < 3; k++) {
        if (origin) origin[k] = c.mc.origin[k];
        if (bin_size) bin_size[k] = c.mc.bin[k];
        if (inv_bin_size) inv_bin_size[k] = c.mc.inv[k];
    }
    return 0;
}

int dem_b200_get_stats(dem_b200_system* s, unsigned long long* nsteps, unsigned long long* nrebuilds,
                       unsigned long long* contacts_last_step) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    Ctrl c;
    int rc = read_ctrl(s, &c);
    if (rc)
        return rc;
    if (nsteps) *nsteps = c.nsteps;
    if (nrebuilds) *nrebuilds = c.nrebuilds;
    if (contacts_last_step) *contacts_last_step = c.n_contacts;
    return 0;
}

int dem_b200_get_history(dem_b200_system* s, uint32_t* owner, uint32_t* other, double* disp3, double* duration,
                         double* relvel_init, size_t capacity, size_t* n) {
    if (!s || !s->initialized || !n)
        return DEMB200_EINVAL;
    *n = 0;
    if (s->P.tang_mode != DEMB200_TANG_MULTISTEP)
        return 0;
    CU(cudaSetDevice(s->cfg.device));
    Ctrl c;
    int rc = read_ctrl(s, &c);
    if (rc)
        return rc;
    const size_t N = s->P.N, Np = s->P.Np, Kn = s->P.Kn, HS = Kn + (size_t)s->P.nW;
    std::vector<VelRec> vr(N);
    std::vector<double4> vals(HS * Np);
    std::vector<uint32_t> nl(Kn * Np);
    std::vector<double> rels;
    CU(cudaMemcpy(vr.data(), s->B.vel[c.cur], N * sizeof(VelRec), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(vals.data(), s->B.hist, HS * Np * sizeof(double4), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(nl.data(), s->B.nl, Kn * Np * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (s->use_hrel) {
        rels.resize(HS * Np);
        CU(cudaMemcpy(rels.data(), s->B.hrel, HS * Np * sizeof(double), cudaMemcpyDeviceToHost));
    }
    // Both partners of a sphere-sphere contact hold a copy; report the one of the higher shape id, which is where
    // Chrono::Multicore keeps it (ChIterativeSolverMulticoreSMC.cpp:194-199).
    size_t cnt = 0;
    auto emit = [&](uint32_t me, uint32_t key, size_t at) {
        if (cnt < capacity) {
            const double4 v = vals[at];
            if (owner) owner[cnt] = me;
            if (other) other[cnt] = key;
            if (disp3) { disp3[3 * cnt] = v.x; disp3[3 * cnt + 1] = v.y; disp3[3 * cnt + 2] = v.z; }
            if (duration) duration[cnt] = v.w * s->P.dt;  // 4th component = steps in contact
            if (relvel_init) relvel_init[cnt] = s->use_hrel ? rels[at] : 0.0;
        }
        cnt++;
    };
    for (size_t i = 0; i < N; i++) {
        const uint32_t me = s->P.shape_base + vr[i].sid;
        const unsigned wmask = (vr[i].meta >> 8) & 0xFFFFu;
        for (int w = 0; w < s->P.nW; w++)
            if ((wmask >> w) & 1u)
                emit(me, (uint32_t)w, (Kn + (size_t)w) * Np + i);
        for (size_t k = 0; k < Kn; k++) {
            if (!((vr[i].amask >> k) & 1ull))
                continue;
            const uint32_t jo = nl[k * Np + i] & ~kEntryFlags;
            const uint32_t key = (jo & kTriFlag) ? (uint32_t)s->P.nW + (jo & ~kTriFlag) : s->P.shape_base + vr[jo].sid;
            if (key < me)
                emit(me, key, k * Np + i);
        }
    }
    *n = cnt;
    return (cnt > capacity && (owner || other || disp3)) ? DEMB200_ECAPACITY : 0;
}

int dem_b200_add_history(dem_b200_system* s, uint32_t owner_shape, uint32_t other_shape, const double disp[3],
                         double duration, double relvel_init) {
    if (!s || !disp)
        return DEMB200_EINVAL;
    if (s->initialized) {
        s->err = "add_history must precede initialize";
        return DEMB200_EINVAL;
    }
    dem_b200_system::HRow r{owner_shape, other_shape, {disp[0], disp[1], disp[2]}, duration, relvel_init};
    s->h_hist.push_back(r);
    return 0;
}

// =============================================================================================
// slab decomposition (one process per GPU): the engine side of the neighbour exchange.  The caller (chrono_b200/slab.py)
// owns the communication buffers and the transport (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
// =============================================================================================
int dem_b200_set_stream(dem_b200_system* s, void* cuda_stream) {
    if (!s || s->initialized)
        return DEMB200_EINVAL;
    if (s->stream && s->own_stream)
        cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
    return 0;
}

int dem_b200_set_sphere_ids(dem_b200_system* s, const uint32_t* ids) {
    if (!s || !ids || s->initialized || s->h_rad.empty())
        return DEMB200_EINVAL;
    s->h_ids.assign(ids, ids + s->h_rad.size());
    return 0;
}

int dem_b200_mgpu_enable(dem_b200_system* s, size_t capacity, double rmax_global) {
    if (!s || s->initialized || capacity == 0 || capacity >= 0xFFFF0000ull)
        return DEMB200_EINVAL;
    s->mgpu = true;
    s->mg_cap = capacity;
    s->mg_rmax = rmax_global;
    refresh_params(s);
    return 0;
}

int dem_b200_mgpu_sizes(dem_b200_system* s, size_t* halo_bytes, size_t* ghost_bytes, size_t* migrant_bytes, double* cut) {
    if (!s || !s->mgpu)
        return DEMB200_EINVAL;
    if (halo_bytes) *halo_bytes = kHaloDoubles * sizeof(double);
    if (ghost_bytes) *ghost_bytes = kGhostDoubles * sizeof(double);
    if (migrant_bytes) *migrant_bytes = (size_t)migrant_doubles(s->P.K) * sizeof(double);
    if (cut) *cut = 2.0 * s->P.rmax + s->P.skin;
    return 0;
}

static int read_slab(dem_b200_system* s, SlabDev* out) {
    CU(cudaMemcpyAsync(s->h_pin, s->B.slab, sizeof(SlabDev), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(out, s->h_pin, sizeof(SlabDev));
    return 0;
}

int dem_b200_mgpu_extract(dem_b200_system* s, double lo, double hi, void* out_left_dev, void* out_right_dev,
                          size_t cap_records, size_t* n_keep, size_t* n_left, size_t* n_right) {
    if (!s || !s->initialized || !s->mgpu || s->mg_phase != 0)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemsetAsync(s->B.slab, 0, sizeof(SlabDev), s->stream));
    const unsigned N = s->P.N;
    k_mgpu_extract<<<(N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, lo, hi, (double*)out_left_dev, (double*)out_right_dev,
                                                            (unsigned)cap_records);
    CU(cudaGetLastError());
    SlabDev sd;
    int rc = read_slab(s, &sd);
    if (rc)
        return rc;
    if (sd.n_out[0] > cap_records || sd.n_out[1] > cap_records) {
        s->err = "mgpu_extract: migration buffer too small";
        return DEMB200_ECAPACITY;
    }
    s->mg_n_own = sd.n_keep;
    s->mg_ng[0] = s->mg_ng[1] = s->mg_ns[0] = s->mg_ns[1] = 0;
    s->mg_phase = 1;
    if (n_keep) *n_keep = sd.n_keep;
    if (n_left) *n_left = sd.n_out[0];
    if (n_right) *n_right = sd.n_out[1];
    return check_device_error(s);
}

int dem_b200_mgpu_append(dem_b200_system* s, const void* in_dev, size_t n, int ghost, int dir) {
    if (!s || !s->initialized || !s->mgpu || (n && !in_dev) || dir < 0 || dir > 1)
        return DEMB200_EINVAL;
    if ((!ghost && s->mg_phase != 1) || (ghost && s->mg_phase != 2)) {
        s->err = "mgpu_append: migrants go between extract and select_ghosts, ghosts between select_ghosts and finish";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    unsigned base;
    if (!ghost) {
        base = s->mg_n_own;
    } else {
        if (dir == 0 && s->mg_ng[1] != 0) {
            s->err = "mgpu_append: append the ghosts of the left neighbour before those of the right one";
            return DEMB200_EINVAL;
        }
        base = s->mg_n_own + s->mg_ng[0] + s->mg_ng[1];
    }
    if ((size_t)base + n > s->mg_cap) {
        s->err = "mgpu_append: local capacity exceeded";
        return DEMB200_ECAPACITY;
    }
    if (n) {
        k_mgpu_append<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->P, s->B, (const double*)in_dev, (unsigned)n, base, ghost);
        CU(cudaGetLastError());
    }
    if (!ghost)
        s->mg_n_own += (unsigned)n;
    else
        s->mg_ng[dir] += (unsigned)n;
    return 0;
}

int dem_b200_mgpu_select_ghosts(dem_b200_system* s, double lo, double hi, double cut, void* out_left_dev, void* out_right_dev,
                                size_t cap_records, size_t* n_left, size_t* n_right) {
    if (!s || !s->initialized || !s->mgpu || s->mg_phase != 1)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    if (!(cut >= 0))
        cut = 2.0 * s->P.rmax + s->P.skin;
    const unsigned n = s->mg_n_own;
    if (n) {
        k_mgpu_select_ghosts<<<(n + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, n, lo, hi, cut, (double*)out_left_dev,
                                                                      (double*)out_right_dev, (unsigned)cap_records);
        CU(cudaGetLastError());
    }
    SlabDev sd;
    int rc = read_slab(s, &sd);
    if (rc)
        return rc;
    if (sd.n_gsend[0] > cap_records || sd.n_gsend[1] > cap_records) {
        s->err = "mgpu_select_ghosts: ghost buffer too small";
        return DEMB200_ECAPACITY;
    }
    s->mg_ns[0] = sd.n_gsend[0];
    s->mg_ns[1] = sd.n_gsend[1];
    s->mg_phase = 2;
    if (n_left) *n_left = sd.n_gsend[0];
    if (n_right) *n_right = sd.n_gsend[1];
    return check_device_error(s);
}

int dem_b200_mgpu_finish_rebuild(dem_b200_system* s) {
    if (!s || !s->initialized || !s->mgpu || s->mg_phase != 2)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    k_mgpu_finish<<<1, 32, 0, s->stream>>>(s->B);
    CU(cudaGetLastError());
    s->mg_n_local = s->mg_n_own + s->mg_ng[0] + s->mg_ng[1];
    if (s->mg_n_local == 0) {
        s->err = "mgpu: a slab without spheres is not supported";
        return DEMB200_EINVAL;
    }
    if (s->p2p && (std::max(s->mg_ns[0], s->mg_ns[1]) > s->p2p_records || std::max(s->mg_ng[0], s->mg_ng[1]) > s->p2p_records)) {
        s->err = "p2p: more ghosts than the landing buffers hold";
        return DEMB200_ECAPACITY;
    }
    s->P.N = s->mg_n_local;
    drop_graph(s);
    s->mg_phase = 0;
    s->slab_rebuilds++;
    s->mg_last_interval = s->step_no - s->mg_last_rebuild_step;
    s->mg_last_rebuild_step = s->step_no;
    s->mg_remap_pending = true;
    s->export_valid = false;
    return recompute_bbox(s);
}

int dem_b200_mgpu_counts(dem_b200_system* s, size_t* n_own, size_t* n_ghost_left, size_t* n_ghost_right, size_t* n_send_left,
                         size_t* n_send_right) {
    if (!s || !s->mgpu)
        return DEMB200_EINVAL;
    if (n_own) *n_own = s->mg_n_own;
    if (n_ghost_left) *n_ghost_left = s->mg_ng[0];
    if (n_ghost_right) *n_ghost_right = s->mg_ng[1];
    if (n_send_left) *n_send_left = s->mg_ns[0];
    if (n_send_right) *n_send_right = s->mg_ns[1];
    return 0;
}

int dem_b200_mgpu_pack(dem_b200_system* s, int dir, void* out_dev) {
    if (!s || !s->initialized || !s->mgpu || dir < 0 || dir > 1 || s->mg_remap_pending)
        return DEMB200_EINVAL;
    const unsigned n = s->mg_ns[dir];
    if (n) {
        k_mgpu_pack<<<(n + 255) / 256, 256, 0, s->stream>>>(s->B, dir, n, (double*)out_dev);
        CU(cudaGetLastError());
    }
    return 0;
}

int dem_b200_mgpu_unpack(dem_b200_system* s, int dir, const void* in_dev) {
    if (!s || !s->initialized || !s->mgpu || dir < 0 || dir > 1 || s->mg_remap_pending)
        return DEMB200_EINVAL;
    const unsigned n = s->mg_ng[dir];
    if (n) {
        k_mgpu_unpack<<<(n + 255) / 256, 256, 0, s->stream>>>(s->B, dir, n, (const double*)in_dev);
        CU(cudaGetLastError());
        s->export_valid = false;
    }
    return 0;
}

int dem_b200_mgpu_want_rebuild_ahead(dem_b200_system* s, int* flag_dev, int steps_ahead) {
    if (!s || !s->initialized || !s->mgpu || !flag_dev || steps_ahead < 0)
        return DEMB200_EINVAL;
    k_mgpu_want<<<1, 32, 0, s->stream>>>(s->P, s->B, flag_dev, steps_ahead);
    CU(cudaGetLastError());
    return 0;
}
int dem_b200_mgpu_want_rebuild(dem_b200_system* s, int* flag_dev) { return dem_b200_mgpu_want_rebuild_ahead(s, flag_dev, 0); }

// ---- direct peer-to-peer halo -----------------------------------------------------------------------------------
// region layout: control block | halo landing [2 sides][2 parities] | migrant landing [2 sides] | ghost landing [2 sides]
// migrants per side and rebuild: a small fraction of the ghost layer in a settling bed, several times the usual in a colliding
// flow (two streams meeting at a slab face) -- an eighth of the ghost capacity, at least 4096
static size_t p2p_mig_records(size_t records) { return std::max<size_t>(4096, records / 8); }
static size_t p2p_region_bytes(size_t records, int K) {
    return kP2PCtlBytes + sizeof(double) * (4 * records * kHaloDoubles + 2 * p2p_mig_records(records) * (size_t)migrant_doubles(K) +
                                            2 * records * kGhostDoubles);
}

int dem_b200_p2p_export(dem_b200_system* s, size_t max_records, void* handle64) {
    if (!s || !s->initialized || !s->mgpu || !handle64 || max_records == 0 || s->p2p_region)
        return DEMB200_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    static_assert(sizeof(P2PCtl) <= kP2PCtlBytes, "control block");
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMalloc(&s->p2p_region, p2p_region_bytes(max_records, s->P.K)));
    CU(cudaMemset(s->p2p_region, 0, p2p_region_bytes(max_records, s->P.K)));
    CU(cudaDeviceSynchronize());
    s->p2p_records = max_records;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->p2p_region));
    memcpy(handle64, &h, 64);
    return 0;
}

int dem_b200_p2p_import(dem_b200_system* s, int rank, int world, const void* handles, int steps_ahead) {
    if (s && world > kMaxRanks) {
        s->err = "p2p_import: the direct halo holds at most 8 ranks (one NVSwitch box); use the NCCL halo (SlabDriver without enable_p2p) beyond that";
        return DEMB200_EINVAL;
    }
    if (!s || !s->p2p_region || !handles || world < 2 || rank < 0 || rank >= world || steps_ahead < 1)
        return DEMB200_EINVAL;
    if (!s->mg_remap_pending) {
        s->err = "p2p_import: switch to the direct halo right after a slab rebuild, before the next step (the first direct step "
                 "must not wait for a halo that nobody sent)";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    P2PDev& X = s->X;
    memset(&X, 0, sizeof(X));
    auto land_of = [&](void* base, int side, int parity) {
        return reinterpret_cast<double*>((char*)base + kP2PCtlBytes) + (size_t)(2 * side + parity) * s->p2p_records * kHaloDoubles;
    };
    for (int r = 0; r < world; r++) {
        void* base = s->p2p_region;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
            CU(cudaIpcOpenMemHandle(&s->p2p_mapped[r], h, cudaIpcMemLazyEnablePeerAccess));
            base = s->p2p_mapped[r];
        }
        X.peer[r] = reinterpret_cast<P2PCtl*>(base);
    }
    X.self = reinterpret_cast<P2PCtl*>(s->p2p_region);
    for (int side = 0; side < 2; side++)
        for (int par = 0; par < 2; par++) {
            X.land[side][par] = land_of(s->p2p_region, side, par);
            const int nb = rank + (side == 0 ? -1 : 1);
            // what I send to my left neighbour lands in ITS "from the right" buffer, and vice versa
            X.peer_land[side][par] = (nb >= 0 && nb < world) ? land_of((void*)X.peer[nb], side == 0 ? 1 : 0, par) : nullptr;
        }
    {
        const size_t R = s->p2p_records, M = p2p_mig_records(R), md = (size_t)migrant_doubles(s->P.K);
        auto mig_of = [&](void* base, int side) {
            return reinterpret_cast<double*>((char*)base + kP2PCtlBytes) + 4 * R * kHaloDoubles + (size_t)side * M * md;
        };
        auto gho_of = [&](void* base, int side) {
            return reinterpret_cast<double*>((char*)base + kP2PCtlBytes) + 4 * R * kHaloDoubles + 2 * M * md + (size_t)side * R * kGhostDoubles;
        };
        for (int side = 0; side < 2; side++) {
            X.mig_land[side] = mig_of(s->p2p_region, side);
            X.gho_land[side] = gho_of(s->p2p_region, side);
            const int nb = rank + (side == 0 ? -1 : 1);
            const bool have = nb >= 0 && nb < world;
            X.peer_mig[side] = have ? mig_of((void*)X.peer[nb], side == 0 ? 1 : 0) : nullptr;
            X.peer_gho[side] = have ? gho_of((void*)X.peer[nb], side == 0 ? 1 : 0) : nullptr;
        }
        X.cap_mig = (unsigned)M;
        X.cap_gho = (unsigned)R;
    }
    int rc = dev_alloc(s, &X.done, 2);
    if (rc)
        return rc;
    CU(cudaMemset(X.done, 0, 2 * sizeof(unsigned)));
    CU(cudaHostAlloc((void**)&s->h_vote, 8 * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(s->h_vote, 0, 8 * sizeof(unsigned long long));
    CU(cudaHostGetDevicePointer((void**)&X.host_vote, s->h_vote, 0));
    X.rank = rank;
    X.world = world;
    X.ahead = steps_ahead;
    X.first_step = s->step_no + 1;
    CU(cudaDeviceSynchronize());
    s->p2p = true;
    drop_graph(s);
    return 0;
}

// The whole rebuild of the slab (migration + new ghost set) without a collective: records go straight into the
// neighbours' landing buffers, counts and arrival flags with them; one host synchronisation at the end to learn the new
// layout.  Every rank must call it at the same step.  counts = {n_own, ghosts from left, from right, sent as ghosts to
// left, to right, migrated out left, out right}.
int dem_b200_p2p_rebuild(dem_b200_system* s, double lo, double hi, size_t counts[7]) {
    if (!s || !s->initialized || !s->mgpu || !s->p2p || s->mg_phase != 0)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    const unsigned long long seq = ++s->p2p_rebuild_seq;
    s->mg_last_interval = s->step_no - s->mg_last_rebuild_step;
    s->mg_last_rebuild_step = s->step_no;
    cudaStream_t st = s->stream;
    const P2PDev& X = s->X;
    CU(cudaMemsetAsync(s->B.slab, 0, sizeof(SlabDev), st));
    const unsigned N = s->P.N, Np = s->P.Np;
    k_mgpu_extract<<<(N + 255) / 256, 256, 0, st>>>(s->P, s->B, lo, hi, X.peer_mig[0], X.peer_mig[1], X.cap_mig);
    k_p2p_publish<<<1, 32, 0, st>>>(s->B, X, 0, seq);
    const unsigned gm = std::min(64u, (X.cap_mig + 255) / 256), gg = std::min(512u, (X.cap_gho + 255) / 256);
    k_p2p_append<<<gm, 256, 0, st>>>(s->P, s->B, X, 0, 0, seq);
    k_p2p_append<<<gm, 256, 0, st>>>(s->P, s->B, X, 0, 1, seq);
    k_mgpu_select_ghosts<<<(Np + 255) / 256, 256, 0, st>>>(s->P, s->B, 0xFFFFFFFFu, lo, hi, 2.0 * s->P.rmax + s->P.skin, X.peer_gho[0],
                                                         X.peer_gho[1], X.cap_gho);
    k_p2p_publish<<<1, 32, 0, st>>>(s->B, X, 1, seq);
    k_p2p_append<<<gg, 256, 0, st>>>(s->P, s->B, X, 1, 0, seq);
    k_p2p_append<<<gg, 256, 0, st>>>(s->P, s->B, X, 1, 1, seq);
    CU(cudaGetLastError());
    SlabDev sd;
    CU(cudaMemcpyAsync(s->h_pin, s->B.slab, sizeof(SlabDev), cudaMemcpyDeviceToHost, st));
    k_mgpu_finish<<<1, 32, 0, st>>>(s->B);  // after the copy: it clears the counters
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    memcpy(&sd, s->h_pin, sizeof(SlabDev));
    s->mg_n_own = sd.n_own;
    for (int d = 0; d < 2; d++) {
        s->mg_ns[d] = sd.n_gsend[d];
        s->mg_ng[d] = sd.g_in[d];
    }
    s->mg_n_local = s->mg_n_own + s->mg_ng[0] + s->mg_ng[1];
    if (counts) {
        counts[0] = sd.n_own; counts[1] = sd.g_in[0]; counts[2] = sd.g_in[1]; counts[3] = sd.n_gsend[0];
        counts[4] = sd.n_gsend[1]; counts[5] = sd.n_out[0]; counts[6] = sd.n_out[1];
    }
    if (s->mg_n_local == 0 || s->mg_n_local > s->mg_cap || std::max(sd.n_gsend[0], sd.n_gsend[1]) > X.cap_gho ||
        std::max(sd.n_out[0], sd.n_out[1]) > X.cap_mig) {
        char msg[400];
        snprintf(msg, sizeof(msg), "p2p_rebuild: slab empty or a buffer capacity exceeded: local %u (own %u + ghosts %u + %u) of capacity %zu, "
                 "ghost senders %u / %u of %u, migrants out %u / %u of %u", s->mg_n_local, s->mg_n_own, s->mg_ng[0], s->mg_ng[1], s->mg_cap,
                 sd.n_gsend[0], sd.n_gsend[1], X.cap_gho, sd.n_out[0], sd.n_out[1], X.cap_mig);
        s->err = msg;
        return DEMB200_ECAPACITY;
    }
    s->P.N = s->mg_n_local;
    drop_graph(s);
    s->mg_remap_pending = true;
    s->export_valid = false;
    int rc = recompute_bbox(s);
    if (rc)
        return rc;
    return check_device_error(s);
}

// Blocks until the all-rank vote of time step `step` (as numbered by dem_b200_step_count) is known; *flag = 1 if any rank
// wants the candidate lists rebuilt.  The vote of step k becomes available while step k+1 runs.
int dem_b200_p2p_poll_vote(dem_b200_system* s, unsigned long long step, int* flag) {
    if (!s || !s->p2p || !flag || step < s->X.first_step || step + 1 > s->step_no)
        return DEMB200_EINVAL;
    volatile unsigned long long* v = s->h_vote + (step & 7ull);
    for (unsigned long long spins = 0;; spins++) {
        const unsigned long long got = *v;
        if ((got >> 1) == step) {
            *flag = (int)(got & 1ull);
            return 0;
        }
        if ((spins & 0xFFFFFull) == 0xFFFFFull && cudaStreamQuery(s->stream) == cudaSuccess && ((*v) >> 1) != step) {
            s->err = "p2p_poll_vote: the stream is idle and the vote never arrived";
            return DEMB200_EINVAL;
        }
    }
}

unsigned long long dem_b200_step_count(const dem_b200_system* s) { return s ? s->step_no : 0ull; }

int dem_b200_export_owned(dem_b200_system* s, uint32_t* sid, double* pos3, double* vel3, double* omega3, size_t capacity,
                          size_t* n) {
    if (!s || !s->initialized || !sid || !n)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    if (!s->d_count) {
        int rc = dev_alloc(s, &s->d_count, 4);
        if (rc)
            return rc;
    }
    const unsigned N = s->P.N;
    unsigned* d_sid = reinterpret_cast<unsigned*>(s->B.rank);  // scratch of the rebuild, free between steps
    unsigned* d_slot = reinterpret_cast<unsigned*>(s->B.cell);  // likewise: storage slot of every exported record
    CU(cudaMemsetAsync(s->d_count, 0, sizeof(unsigned), s->stream));
    k_export_owned<<<(N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, s->d_count, (unsigned)std::min<size_t>(capacity, N), d_sid, d_slot,
                                                            s->d_pos3, s->d_vel3, s->d_om3);
    CU(cudaGetLastError());
    s->export_valid = false;
    CU(cudaMemcpyAsync(s->h_pin, s->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    unsigned c;
    memcpy(&c, s->h_pin, sizeof(unsigned));
    *n = c;
    if (c > capacity)
        return DEMB200_ECAPACITY;
    s->owned_export_n = c;
    s->owned_export_step = s->step_no;
    s->owned_export_seq = s->p2p_rebuild_seq + s->slab_rebuilds;
    CU(cudaMemcpyAsync(sid, d_sid, c * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    if (pos3) CU(cudaMemcpyAsync(pos3, s->d_pos3, 3 * (size_t)c * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (vel3) CU(cudaMemcpyAsync(vel3, s->d_vel3, 3 * (size_t)c * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (omega3) CU(cudaMemcpyAsync(omega3, s->d_om3, 3 * (size_t)c * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    return check_device_error(s);
}

// The inverse of dem_b200_export_owned for the slab-mode host round trip: the n records of the LAST export (same order), with
// possibly changed values, go back into the spheres they came from.  Only valid while nothing (step, rebuild) has happened
// since that export.  New positions invalidate the candidate lists and the ghost copies held by the neighbours: the caller
// runs the slab rebuild protocol before the next step (chrono_b200/slab.py: SlabDriver.import_owned).
int dem_b200_import_owned(dem_b200_system* s, size_t n, const double* pos3, const double* vel3, const double* omega3) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    if (n != s->owned_export_n || s->owned_export_step != s->step_no || s->owned_export_seq != s->p2p_rebuild_seq + s->slab_rebuilds) {
        s->err = "import_owned: must directly follow dem_b200_export_owned (same records, same order, no step or rebuild in between)";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    const size_t bytes = 3 * n * sizeof(double);
    if (pos3) CU(cudaMemcpyAsync(s->d_pos3, pos3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (vel3) CU(cudaMemcpyAsync(s->d_vel3, vel3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (omega3) CU(cudaMemcpyAsync(s->d_om3, omega3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (n) {
        k_import_owned<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->B, (unsigned)n, reinterpret_cast<unsigned*>(s->B.cell),
                                                                          pos3 ? s->d_pos3 : nullptr, vel3 ? s->d_vel3 : nullptr,
                                                                          omega3 ? s->d_om3 : nullptr);
        CU(cudaGetLastError());
    }
    s->export_valid = false;
    s->accel_src_valid = s->accel_export_valid = false;
    if (pos3 && !s->mgpu) {
        const unsigned one = 1;
        CU(cudaMemcpyAsync(&s->B.ctrl->need_rebuild, &one, sizeof(unsigned), cudaMemcpyHostToDevice, s->stream));
        int rc = recompute_bbox(s);
        if (rc)
            return rc;
    }
    CU(cudaStreamSynchronize(s->stream));  // the host buffers may be reused by the caller
    return 0;
}

// ---- per-contact records: Chrono::Dem SetRecordingContactInfo / getNormalForce ... / WriteContactInfoFile -----------------
int dem_b200_enable_contact_info(dem_b200_system* s, int enable) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    if (s->mgpu) {
        s->err = "contact info recording is not available on a slab engine";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaStreamSynchronize(s->stream));
    if (!enable) {
        s->contact_info = false;
        if (!s->user_recording)
            return dem_b200_enable_recording(s, 0, 0);
        return 0;
    }
    const bool keep_user = s->user_recording;
    int rc = dem_b200_enable_recording(s, 1, s->max_pairs);  // the recording instantiations of the force kernel write the records
    s->user_recording = keep_user;
    if (rc)
        return rc;
    if (!s->B.cinfo) {
        const size_t rows = (size_t)s->P.Kn + (size_t)s->P.nW;
        rc = dev_alloc(s, &s->B.cinfo, rows * s->P.Np * kCInfo);
        if (rc)
            return rc;
        CU(cudaMemset(s->B.cinfo, 0, rows * s->P.Np * kCInfo * sizeof(double)));
    }
    s->contact_info = true;
    return 0;
}

int dem_b200_get_contact_info(dem_b200_system* s, size_t sphere, uint32_t other_shape, double info13[13], int* found) {
    if (!s || !s->initialized || !info13 || !found || sphere >= s->P.N)
        return DEMB200_EINVAL;
    if (!s->contact_info || !s->B.cinfo) {
        s->err = "get_contact_info: call dem_b200_enable_contact_info(s, 1) first";
        return DEMB200_EINVAL;
    }
    if (s->P.tang_mode != DEMB200_TANG_MULTISTEP) {
        s->err = "contact info needs the MultiStep contact map";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    double* d_out = nullptr;
    if (!s->d_cinfo_out) {
        int rc = dev_alloc(s, &s->d_cinfo_out, 16);
        if (rc)
            return rc;
    }
    d_out = s->d_cinfo_out;
    CU(cudaMemsetAsync(d_out, 0, 16 * sizeof(double), s->stream));
    k_find_contact<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, (unsigned)sphere, other_shape, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->h_pin, d_out, 14 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    *found = s->h_pin[0] != 0.0;
    for (int c = 0; c < kCInfo; c++)
        info13[c] = *found ? s->h_pin[1 + c] : 0.0;
    return 0;
}

int dem_b200_get_contact_infos(dem_b200_system* s, uint32_t* bi, uint32_t* bj, double* info13, size_t capacity, size_t* n) {
    if (!s || !s->initialized || !n)
        return DEMB200_EINVAL;
    if (!s->contact_info || !s->B.cinfo || s->P.tang_mode != DEMB200_TANG_MULTISTEP) {
        s->err = "get_contact_infos: needs dem_b200_enable_contact_info(s, 1) and MultiStep friction";
        return DEMB200_EINVAL;
    }
    CU(cudaSetDevice(s->cfg.device));
    if (!s->d_count) {
        int rc = dev_alloc(s, &s->d_count, 4);
        if (rc)
            return rc;
    }
    unsigned *d_bi = nullptr, *d_bj = nullptr;
    double* d_info = nullptr;
    const size_t cap = (bi && bj && info13) ? capacity : 0;
    if (cap) {
        CU(cudaMalloc((void**)&d_bi, cap * sizeof(unsigned)));
        CU(cudaMalloc((void**)&d_bj, cap * sizeof(unsigned)));
        CU(cudaMalloc((void**)&d_info, cap * kCInfo * sizeof(double)));
    }
    CU(cudaMemsetAsync(s->d_count, 0, sizeof(unsigned), s->stream));
    k_export_contacts<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, s->d_count, (unsigned)std::min<size_t>(cap, 0xFFFFFFFFull),
                                                                    d_bi, d_bj, d_info);
    cudaError_t e = cudaGetLastError();
    unsigned c = 0;
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s->h_pin, s->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(s->stream);
    memcpy(&c, s->h_pin, sizeof(unsigned));
    *n = c;
    if (e == cudaSuccess && cap && c <= cap && c) {
        cudaMemcpy(bi, d_bi, c * sizeof(unsigned), cudaMemcpyDeviceToHost);
        cudaMemcpy(bj, d_bj, c * sizeof(unsigned), cudaMemcpyDeviceToHost);
        e = cudaMemcpy(info13, d_info, (size_t)c * kCInfo * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_bi); cudaFree(d_bj); cudaFree(d_info);
    if (e != cudaSuccess) {
        s->err = std::string("get_contact_infos: ") + cudaGetErrorString(e);
        return DEMB200_ECUDA;
    }
    return (cap && c > cap) ? DEMB200_ECAPACITY : 0;
}

// Chrono::Dem SetBCPlaneRotation (ChSystemDem_impl.cpp:993-1001): the wall's surface moves with vel + omega x (x - center) at
// the contact point (ChDemBoundaryConditions.cuh:394); the wall geometry stays put.
int dem_b200_set_wall_rotation(dem_b200_system* s, int w, const double center[3], const double omega[3]) {
    if (!s || w < 0 || w >= s->P.nW || !center || !omega)
        return DEMB200_EINVAL;
    for (int k = 0; k < 3; k++) {
        s->W.w[w].rc[k] = center[k];
        s->W.w[w].omg[k] = omega[k];
    }
    return s->initialized ? upload_walls(s) : 0;
}

int dem_b200_clear_error(dem_b200_system* s) {
    if (!s || !s->initialized)
        return DEMB200_EINVAL;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemsetAsync(&s->B.ctrl->err, 0, sizeof(unsigned), s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->err.clear();
    return 0;
}

}  // extern "C"

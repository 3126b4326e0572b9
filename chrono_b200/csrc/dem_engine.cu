// =============================================================================
// dem_engine.cu -- host side of the B200 SMC-DEM engine and its C ABI (include/chrono_b200_dem.h).
//
// Replaces, for the hot path, ChSystemDem_impl::initializeSpheres / AdvanceSimulation / getters
// (reference: src/chrono_dem/physics/ChSystemDem_impl.cpp:1130-1166, src/chrono_dem/gpu/ChDemSMC.cu:619-691).
// Unlike the reference there is no managed memory, no device synchronisation between kernels and no host
// round trip inside a step: the 9 launches of a step are enqueued on one stream and replayed from a CUDA graph
// holding two steps (the contact-history double buffer has period two).
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/chrono_b200_dem.h"
#include "dem_kernels.cuh"

using namespace demb200;

namespace {
thread_local std::string g_create_error;
}

struct dem_b200_system {
    dem_b200_config cfg{};
    Params P{};
    Buffers B{};
    // scene staged on the host until initialize()
    std::vector<double> h_pos, h_vel, h_om, h_rad;
    std::vector<uint8_t> h_fixed;
    bool any_fixed = false;
    bool initialized = false;
    cudaStream_t stream = nullptr;
    unsigned ncell = 0, ntiles = 0;
    cudaGraphExec_t graph2 = nullptr;  // two steps
    bool recording = false;
    size_t max_pairs = 0;
    bool use_hrel = false;
    double time = 0.0;
    std::string err;
    // scratch (device, by user index) and pinned host staging
    double* d_pos3 = nullptr; double* d_vel3 = nullptr; double* d_om3 = nullptr;
    double* d_red = nullptr;            // reduction scratch (2 x 8 bytes)
    double* h_pin = nullptr;            // pinned, 16 doubles
    bool export_valid = false;
    std::vector<void*> allocs;
    // history staged before initialize (add_history)
    struct HRow { uint32_t owner, other; double d[3], dur, rel; };
    std::vector<HRow> h_hist;
};

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
    P.K = c.history_slots > 0 ? c.history_slots : 12;
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
    // union of the box-wall AABBs
    P.has_wall_bb = 0;
    for (int w = 0; w < P.nW; w++) {
        Wall& W = P.walls[w];
        if (W.type != WALL_BOX)
            continue;
        double ext[3];
        abs_rotate(W.rot, W.hdims, ext);  // ComputeAABBBox, ChCollisionSystemMulticore.cpp:395-406 (envelope 0)
        for (int k = 0; k < 3; k++) {
            W.amin[k] = W.pos[k] - ext[k];
            W.amax[k] = W.pos[k] + ext[k];
            if (!P.has_wall_bb) {
                P.wall_bb_min[k] = W.amin[k];
                P.wall_bb_max[k] = W.amax[k];
            } else {
                P.wall_bb_min[k] = std::min(P.wall_bb_min[k], W.amin[k]);
                P.wall_bb_max[k] = std::max(P.wall_bb_max[k], W.amax[k]);
            }
        }
        P.has_wall_bb = 1;
    }
    P.shape_base = (unsigned)P.nW;
}

bool need_roll(const dem_b200_system* s) {
    for (int k = 0; k < 2; k++)
        if (s->P.comp[k].mu_roll > 0 || s->P.comp[k].mu_spin > 0)
            return true;
    return false;
}

template <bool REC>
void launch_force(dem_b200_system* s, const Buffers& B, unsigned blocks) {
    const Params& P = s->P;
    const bool hist = (P.tang_mode == DEMB200_TANG_MULTISTEP);
    const bool roll = need_roll(s);
    if (hist) {
        if (roll) k_force_integrate<true, true, REC><<<blocks, kForceThreads, 0, s->stream>>>(P, B);
        else k_force_integrate<true, false, REC><<<blocks, kForceThreads, 0, s->stream>>>(P, B);
    } else {
        if (roll) k_force_integrate<false, true, REC><<<blocks, kForceThreads, 0, s->stream>>>(P, B);
        else k_force_integrate<false, false, REC><<<blocks, kForceThreads, 0, s->stream>>>(P, B);
    }
}

constexpr int kNumKernels = 9;
const char* kKernelNames[kNumKernels] = {"k_grid_update", "memset_bin_count", "k_bin_count", "k_scan_tile_sums",
                                         "k_scan_sums", "k_scan_apply", "k_scatter_perm", "k_gather_sorted",
                                         "k_force_integrate"};

// Enqueue one step.  ev: optional kNumKernels+1 events recorded around each launch (profiling).
int enqueue_step(dem_b200_system* s, cudaEvent_t* ev) {
    const Params& P = s->P;
    Buffers& B = s->B;
    const unsigned N = P.N;
    const unsigned nb256 = (N + 255) / 256;
    cudaStream_t st = s->stream;
    int k = 0;
    auto mark = [&]() {
        if (ev)
            cudaEventRecord(ev[k++], st);
    };
    mark();
    k_grid_update<<<1, 32, 0, st>>>(P, B);
    mark();
    CU(cudaMemsetAsync(B.cell_count, 0, sizeof(uint32_t) * s->ncell, st));
    if (s->recording) {
        CU(cudaMemsetAsync(B.pair_count, 0, sizeof(unsigned long long), st));
        CU(cudaMemsetAsync(B.n_contacts, 0, sizeof(unsigned long long), st));
    }
    mark();
    if (s->recording)
        k_bin_count<true><<<nb256, 256, 0, st>>>(P, B);
    else
        k_bin_count<false><<<nb256, 256, 0, st>>>(P, B);
    mark();
    k_scan_tile_sums<<<s->ntiles, kScanThreads, 0, st>>>(s->ncell, B.cell_count, B.block_sums);
    mark();
    k_scan_sums<<<1, kScanThreads, 0, st>>>(s->ntiles, B.block_sums);
    mark();
    k_scan_apply<<<s->ntiles, kScanThreads, 0, st>>>(s->ncell, N, B.cell_count, B.block_sums, B.cell_start);
    mark();
    k_scatter_perm<<<nb256, 256, 0, st>>>(P, B);
    mark();
    if (P.integrator == DEMB200_CHUNG)
        k_gather_sorted<true><<<nb256, 256, 0, st>>>(P, B);
    else
        k_gather_sorted<false><<<nb256, 256, 0, st>>>(P, B);
    mark();
    const unsigned fb = (N + kForceThreads - 1) / kForceThreads;
    if (s->recording)
        launch_force<true>(s, B, fb);
    else
        launch_force<false>(s, B, fb);
    mark();
    CU(cudaGetLastError());
    // the history written this step is next step's input
    std::swap(B.hkey_old, B.hkey_new);
    std::swap(B.hval_old, B.hval_new);
    std::swap(B.hrel_old, B.hrel_new);
    s->time += P.dt;
    s->export_valid = false;
    return 0;
}

void drop_graph(dem_b200_system* s) {
    if (s->graph2) {
        cudaGraphExecDestroy(s->graph2);
        s->graph2 = nullptr;
    }
}

int build_graph(dem_b200_system* s) {
    cudaGraph_t g = nullptr;
    CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    double t = s->time;
    int rc = enqueue_step(s, nullptr);
    if (rc == 0)
        rc = enqueue_step(s, nullptr);
    s->time = t;  // capturing does not advance time
    cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    if (rc)
        return rc;
    if (e != cudaSuccess) {
        s->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e);
        return DEMB200_ECUDA;
    }
    CU(cudaGraphInstantiate(&s->graph2, g, 0));
    cudaGraphDestroy(g);
    return 0;
}

int check_device_error(dem_b200_system* s) {
    unsigned e = 0;
    CU(cudaMemcpyAsync(s->h_pin, s->B.err, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(&e, s->h_pin, sizeof(unsigned));
    if (!e)
        return 0;
    if (e & ERR_NAN) { s->err = "non-finite sphere state"; return DEMB200_ENAN; }
    if (e & ERR_GRID_BIN_TOO_SMALL) { s->err = "broadphase bin edge smaller than the largest sphere diameter"; return DEMB200_EGRID; }
    if (e & ERR_GRID_OUT_OF_RANGE) { s->err = "sphere outside the broadphase grid"; return DEMB200_EGRID; }
    if (e & (ERR_HISTORY_OVERFLOW | ERR_CONTACT_LIST_OVERFLOW)) { s->err = "contact history slots exhausted"; return DEMB200_EHISTORY; }
    if (e & ERR_PAIR_CAPACITY) { s->err = "pair recording buffer overflow"; return DEMB200_ECAPACITY; }
    s->err = "unknown device error";
    return DEMB200_ECUDA;
}

int recompute_bbox(dem_b200_system* s) {
    unsigned long long init[6];
    for (int k = 0; k < 3; k++) {
        init[k] = s->P.has_wall_bb ? enc_ord_h(s->P.wall_bb_min[k]) : enc_ord_h(INFINITY);
        init[3 + k] = s->P.has_wall_bb ? enc_ord_h(s->P.wall_bb_max[k]) : enc_ord_h(-INFINITY);
    }
    CU(cudaMemcpyAsync(s->B.bbox, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));  // init[] is on the stack
    k_bbox_reduce<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P.N, s->B.posA, s->B.bbox);
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
    if (!s->recording && nsteps >= 2) {
        if (!s->graph2) {
            int rc = build_graph(s);
            if (rc)
                return rc;
        }
        for (; done + 2 <= nsteps; done += 2) {
            CU(cudaGraphLaunch(s->graph2, s->stream));
            s->time += 2 * s->P.dt;
        }
        s->export_valid = false;
    }
    for (; done < nsteps; done++) {
        int rc = enqueue_step(s, nullptr);
        if (rc)
            return rc;
        // an odd direct step flips the history buffers relative to the captured graph
        drop_graph(s);
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
    if (s->stream)
        cudaStreamDestroy(s->stream);
    delete s;
}

const char* dem_b200_last_error(const dem_b200_system* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int dem_b200_set_config(dem_b200_system* s, const dem_b200_config* cfg) {
    if (!s || !cfg)
        return DEMB200_EINVAL;
    if (s->initialized) {
        const dem_b200_config& o = s->cfg;
        const int Kold = o.history_slots > 0 ? o.history_slots : 12, Knew = cfg->history_slots > 0 ? cfg->history_slots : 12;
        if (Kold != Knew || o.bins_per_axis[0] != cfg->bins_per_axis[0] || o.bins_per_axis[1] != cfg->bins_per_axis[1] ||
            o.bins_per_axis[2] != cfg->bins_per_axis[2] || o.device != cfg->device ||
            (o.integrator != cfg->integrator && (o.integrator == DEMB200_CHUNG || cfg->integrator == DEMB200_CHUNG)) ||
            o.tangential_mode != cfg->tangential_mode) {
            s->err = "set_config: history_slots, bins_per_axis, device, tangential_mode and (to/from) Chung cannot change after initialize";
            return DEMB200_EINVAL;
        }
    }
    s->cfg = *cfg;
    refresh_params(s);
    drop_graph(s);
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
    Wall& W = s->P.walls[s->P.nW];
    memset(&W, 0, sizeof(W));
    W.type = type;
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
int dem_b200_set_wall_velocity(dem_b200_system* s, int w, const double pos[3], const double vel[3]) {
    if (!s || w < 0 || w >= s->P.nW)
        return DEMB200_EINVAL;
    for (int k = 0; k < 3; k++) {
        if (pos) s->P.walls[w].pos[k] = pos[k];
        if (vel) s->P.walls[w].vel[k] = vel[k];
    }
    refresh_params(s);
    drop_graph(s);
    return 0;
}
int dem_b200_num_walls(const dem_b200_system* s) { return s ? s->P.nW : 0; }

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
    refresh_params(s);
    Params& P = s->P;
    Buffers& B = s->B;
    P.N = (unsigned)n;
    const long long nc = (long long)P.bins[0] * P.bins[1] * P.bins[2];
    if (P.bins[0] < 1 || P.bins[1] < 1 || P.bins[2] < 1 || nc >= (1ll << 31)) {
        s->err = "bad bins_per_axis";
        return DEMB200_EINVAL;
    }
    s->ncell = (unsigned)nc;
    s->ntiles = (s->ncell + kScanTile - 1) / kScanTile;
    P.rmax = *std::max_element(s->h_rad.begin(), s->h_rad.end());
    const size_t K = (size_t)P.K;
    const bool hist = (P.tang_mode == DEMB200_TANG_MULTISTEP);
    s->use_hrel = hist && P.use_mat_props && (P.force_model == DEMB200_HOOKE || P.force_model == DEMB200_FLORES);

    int rc = 0;
    rc |= dev_alloc(s, &B.posA, n); rc |= dev_alloc(s, &B.posB, n);
    rc |= dev_alloc(s, &B.velA, 6 * n); rc |= dev_alloc(s, &B.velB, 6 * n);
    rc |= dev_alloc(s, &B.sidA, n); rc |= dev_alloc(s, &B.sidB, n);
    if (P.integrator == DEMB200_CHUNG) {
        rc |= dev_alloc(s, &B.accA, 6 * n); rc |= dev_alloc(s, &B.accB, 6 * n);
    }
    rc |= dev_alloc(s, &B.cell, n); rc |= dev_alloc(s, &B.rank, n); rc |= dev_alloc(s, &B.perm, n);
    rc |= dev_alloc(s, &B.cell_count, (size_t)s->ncell + 8); rc |= dev_alloc(s, &B.cell_start, (size_t)s->ncell + 8);
    rc |= dev_alloc(s, &B.block_sums, (size_t)s->ntiles + 8);
    rc |= dev_alloc(s, &B.grid, 1); rc |= dev_alloc(s, &B.bbox, 8); rc |= dev_alloc(s, &B.err, 4);
    rc |= dev_alloc(s, &B.n_contacts, 2); rc |= dev_alloc(s, &B.pair_count, 2);
    if (hist) {
        rc |= dev_alloc(s, &B.hkey_old, n * K); rc |= dev_alloc(s, &B.hkey_new, n * K);
        rc |= dev_alloc(s, &B.hval_old, n * K); rc |= dev_alloc(s, &B.hval_new, n * K);
        if (s->use_hrel) {
            rc |= dev_alloc(s, &B.hrel_old, n * K); rc |= dev_alloc(s, &B.hrel_new, n * K);
        }
    }
    if (s->any_fixed)
        rc |= dev_alloc(s, &B.flags, n);
    rc |= dev_alloc(s, &s->d_pos3, 3 * n); rc |= dev_alloc(s, &s->d_vel3, 3 * n); rc |= dev_alloc(s, &s->d_om3, 3 * n);
    rc |= dev_alloc(s, &s->d_red, 4);
    if (rc)
        return DEMB200_ECUDA;

    // upload (storage order = user order initially; the first step sorts by bin)
    {
        std::vector<double4> hp(n);
        std::vector<double> hv(6 * n);
        std::vector<uint32_t> hs(n);
        for (size_t i = 0; i < n; i++) {
            hp[i] = make_double4(s->h_pos[3 * i], s->h_pos[3 * i + 1], s->h_pos[3 * i + 2], s->h_rad[i]);
            for (int k = 0; k < 3; k++) {
                hv[6 * i + k] = s->h_vel[3 * i + k];
                hv[6 * i + 3 + k] = s->h_om[3 * i + k];
            }
            hs[i] = (uint32_t)i;
        }
        CU(cudaMemcpy(B.posA, hp.data(), n * sizeof(double4), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(B.velA, hv.data(), 6 * n * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(B.sidA, hs.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (B.accA) {
            CU(cudaMemset(B.accA, 0, 6 * n * sizeof(double)));
            CU(cudaMemset(B.accB, 0, 6 * n * sizeof(double)));
        }
        if (B.flags)
            CU(cudaMemcpy(B.flags, s->h_fixed.data(), n, cudaMemcpyHostToDevice));
    }
    CU(cudaMemset(B.err, 0, 4 * sizeof(unsigned)));
    if (hist) {
        CU(cudaMemset(B.hkey_old, 0xFF, n * K * sizeof(uint32_t)));
        CU(cudaMemset(B.hkey_new, 0xFF, n * K * sizeof(uint32_t)));
        CU(cudaMemset(B.hval_old, 0, n * K * sizeof(double4)));
        CU(cudaMemset(B.hval_new, 0, n * K * sizeof(double4)));
        if (s->use_hrel) {
            CU(cudaMemset(B.hrel_old, 0, n * K * sizeof(double)));
            CU(cudaMemset(B.hrel_new, 0, n * K * sizeof(double)));
        }
        // history supplied before initialize (checkpoint restart)
        if (!s->h_hist.empty()) {
            std::vector<uint32_t> keys(n * K, kEmptyKey);
            std::vector<double4> vals(n * K, make_double4(0, 0, 0, 0));
            std::vector<double> rels(n * K, 0.0);
            std::vector<int> fill(n, 0);
            for (auto& r : s->h_hist) {
                if (r.owner < P.shape_base || r.owner - P.shape_base >= n) {
                    s->err = "add_history: owner is not a sphere shape";
                    return DEMB200_EINVAL;
                }
                size_t sid = r.owner - P.shape_base;
                if (fill[sid] >= (int)K) {
                    s->err = "add_history: too many rows for one sphere";
                    return DEMB200_EHISTORY;
                }
                size_t at = sid * K + fill[sid]++;
                keys[at] = r.other;
                vals[at] = make_double4(r.d[0], r.d[1], r.d[2], r.dur);
                rels[at] = r.rel;
            }
            CU(cudaMemcpy(B.hkey_old, keys.data(), n * K * sizeof(uint32_t), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(B.hval_old, vals.data(), n * K * sizeof(double4), cudaMemcpyHostToDevice));
            if (s->use_hrel)
                CU(cudaMemcpy(B.hrel_old, rels.data(), n * K * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    s->h_pos.clear(); s->h_pos.shrink_to_fit();
    s->h_vel.clear(); s->h_vel.shrink_to_fit();
    s->h_om.clear(); s->h_om.shrink_to_fit();
    s->initialized = true;
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
    CU(cudaSetDevice(s->cfg.device));
    const size_t bytes = 3 * (size_t)s->P.N * sizeof(double);
    if (pos3) CU(cudaMemcpyAsync(s->d_pos3, pos3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (vel3) CU(cudaMemcpyAsync(s->d_vel3, vel3, bytes, cudaMemcpyHostToDevice, s->stream));
    if (omega3) CU(cudaMemcpyAsync(s->d_om3, omega3, bytes, cudaMemcpyHostToDevice, s->stream));
    k_import_state<<<(s->P.N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, pos3 ? s->d_pos3 : nullptr,
                                                                vel3 ? s->d_vel3 : nullptr, omega3 ? s->d_om3 : nullptr);
    CU(cudaGetLastError());
    s->export_valid = false;
    if (pos3)
        return recompute_bbox(s);
    return 0;
}

int dem_b200_advance_host(dem_b200_system* s, size_t n, const double* pos3_in, const double* vel3_in,
                          const double* omega3_in, int nsteps, double* pos3_out, double* vel3_out,
                          double* omega3_out) {
    if (!s || !s->initialized || n != s->P.N)
        return DEMB200_EINVAL;
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
    if (which == DEMB200_RED_NUM_CONTACTS) {
        if (s->P.tang_mode != DEMB200_TANG_MULTISTEP) {
            s->err = "NUM_CONTACTS needs MultiStep history (or use recording)";
            return DEMB200_EINVAL;
        }
        CU(cudaMemsetAsync(s->d_red, 0, 16, s->stream));
        unsigned long long tot = (unsigned long long)N * s->P.K;
        k_count_history<<<(unsigned)((tot + 255) / 256), 256, 0, s->stream>>>(tot, s->B.hkey_old,
                                                                               (unsigned long long*)s->d_red);
        CU(cudaMemcpyAsync(s->h_pin, s->d_red, 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        unsigned long long c;
        memcpy(&c, s->h_pin, 8);
        *out = (double)c;
        return 0;
    }
    if (which < 0 || which > 5)
        return DEMB200_EINVAL;
    unsigned long long init[2] = {0ull, enc_ord_h(-INFINITY)};
    memcpy(s->h_pin + 12, init, 16);
    CU(cudaMemcpyAsync(s->d_red, s->h_pin + 12, 16, cudaMemcpyHostToDevice, s->stream));
    k_reduce<<<(N + 255) / 256, 256, 0, s->stream>>>(s->P, s->B, which, arg, s->d_red,
                                                      (unsigned long long*)(s->d_red + 1));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->h_pin, s->d_red, 16, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (which == 2 || which == 4 || which == 5) {
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
    CU(cudaSetDevice(s->cfg.device));
    drop_graph(s);
    s->recording = enable != 0;
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
    CU(cudaMemcpyAsync(s->h_pin, s->B.pair_count, 8, cudaMemcpyDeviceToHost, s->stream));
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
    GridDev g;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(&g, s->B.grid, sizeof(GridDev), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) {
        if (origin) origin[k] = g.origin[k];
        if (bin_size) bin_size[k] = g.bin[k];
        if (inv_bin_size) inv_bin_size[k] = g.inv[k];
    }
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
    CU(cudaStreamSynchronize(s->stream));
    const size_t N = s->P.N, K = s->P.K;
    std::vector<uint32_t> keys(N * K);
    std::vector<double4> vals(N * K);
    std::vector<double> rels;
    CU(cudaMemcpy(keys.data(), s->B.hkey_old, N * K * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(vals.data(), s->B.hval_old, N * K * sizeof(double4), cudaMemcpyDeviceToHost));
    if (s->use_hrel) {
        rels.resize(N * K);
        CU(cudaMemcpy(rels.data(), s->B.hrel_old, N * K * sizeof(double), cudaMemcpyDeviceToHost));
    }
    size_t c = 0;
    for (size_t i = 0; i < N; i++)
        for (size_t k = 0; k < K; k++) {
            if (keys[i * K + k] == kEmptyKey)
                continue;
            if (c < capacity) {
                if (owner) owner[c] = s->P.shape_base + (uint32_t)i;
                if (other) other[c] = keys[i * K + k];
                if (disp3) { disp3[3 * c] = vals[i * K + k].x; disp3[3 * c + 1] = vals[i * K + k].y; disp3[3 * c + 2] = vals[i * K + k].z; }
                if (duration) duration[c] = vals[i * K + k].w;
                if (relvel_init) relvel_init[c] = s->use_hrel ? rels[i * K + k] : 0.0;
            }
            c++;
        }
    *n = c;
    return (c > capacity && (owner || other || disp3)) ? DEMB200_ECAPACITY : 0;
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

}  // extern "C"

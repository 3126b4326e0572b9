// =============================================================================
// dem_types.h -- shared host/device declarations of the B200 SMC-DEM engine.
//
// HBM layout (fp64, one record per sphere, two copies "A"/"B" that ping-pong every step):
//   posr : double4 (x, y, z, radius)            32 B, one DRAM sector per neighbour gather
//   velw : double[6] (vx,vy,vz, wx,wy,wz)       48 B, gathered only for spheres actually in contact
//   sid  : uint32  stable sphere id (= user index); shape id = shape_base + sid (Multicore numbering)
// "A" is the state at the start of a step in last step's cell order; "B" is the same state re-sorted by
// broadphase cell (z-major Multicore hash, src/chrono/collision/multicore/ChCollisionUtils.h:63-65).
// Contact history (tangential displacement map keyed by contact pair) lives in rows of K slots indexed by the
// STABLE id of the owning sphere (owner = higher shape id, as ChIterativeSolverMulticoreSMC.cpp:194-199), so it
// never moves when spheres are re-sorted:  hkey: uint32[K] partner shape id, hval: double4[K] (disp.xyz, duration).
// =============================================================================
#pragma once
#include <cstdint>

namespace demb200 {

constexpr int kMaxWalls = 16;
constexpr unsigned kEmptyKey = 0xFFFFFFFFu;
constexpr int kMaxContactsPerSphere = 24;  // sphere-sphere contacts a thread can stage (monodisperse max is 12)

// device error bits (dem_b200 error codes are derived from these at sync points)
enum : unsigned {
    ERR_GRID_BIN_TOO_SMALL = 1u,
    ERR_GRID_OUT_OF_RANGE = 2u,
    ERR_HISTORY_OVERFLOW = 4u,
    ERR_CONTACT_LIST_OVERFLOW = 8u,
    ERR_NAN = 16u,
    ERR_PAIR_CAPACITY = 32u
};

enum WallType : int { WALL_BOX = 0, WALL_PLANE = 1 };

// Composite material of a contact class, widened from the float values the reference computes
// (src/chrono/physics/ChContactMaterialSMC.cpp:107-130).
struct Comp {
    double E_eff, G_eff, mu, mu_roll, mu_spin, cr, adh, adh_dmt, adh_perko, kn, kt, gn, gt;
    // hoisted, material-only factors of the Hertz / PlainCoulomb / Flores damping (ChIterativeSolverMulticoreSMC.cpp:283-287)
    double hertz_damp;  // -2*sqrt(5/6)*beta
};

struct Wall {
    int type;
    double pos[3];     // box centre / plane point (world)
    double rot[4];     // box orientation (w,x,y,z)
    double hdims[3];   // box half dimensions / plane unit normal
    double vel[3];     // velocity of the wall body (moving boundaries)
    double amin[3], amax[3];  // world AABB (ChCollisionSystemMulticore.cpp:395-406)
};

struct Params {
    unsigned N;
    int nW;
    int K;
    int force_model, adhesion_model, tang_mode, use_mat_props, integrator;
    double char_vel, min_slip, min_roll, min_spin, dt;
    double g[3];
    double mass_coef, wall_mass;
    Comp comp[3];
    Wall walls[kMaxWalls];
    int bins[3];
    unsigned shape_base;  // shape id of sphere sid is shape_base + sid
    double rmax;          // largest sphere radius (grid validity check)
    double wall_bb_min[3], wall_bb_max[3];  // union of wall AABBs (infinite planes excluded)
    int has_wall_bb;
};

// Broadphase grid of the current step (device memory, written by k_grid_update)
struct GridDev {
    double origin[3], bin[3], inv[3];
    double wmin[kMaxWalls][3], wmax[kMaxWalls][3];  // wall AABBs offset by the grid origin (ChBroadphase.cpp:168-176)
};

struct Buffers {
    // state
    double4* posA; double4* posB;
    double* velA; double* velB;
    uint32_t* sidA; uint32_t* sidB;
    double* accA; double* accB;      // previous-step acceleration (Chung only), 6 per sphere
    uint8_t* flags;                  // by sid: bit0 = fixed; may be null
    // broadphase
    uint32_t* cell; uint32_t* rank; uint32_t* perm;
    uint32_t* cell_count; uint32_t* cell_start; uint32_t* block_sums;
    GridDev* grid;
    unsigned long long* bbox;        // 6 order-preserving encoded doubles: min xyz, max xyz of sphere AABBs
    // history (rows by sid)
    uint32_t* hkey_old; uint32_t* hkey_new;
    double4* hval_old; double4* hval_new;
    double* hrel_old; double* hrel_new;  // initial normal speed (only Hooke/Flores with material properties)
    // diagnostics / recording
    unsigned* err;
    unsigned long long* n_contacts;  // running count of sphere-sphere + sphere-wall contacts of the last step
    double* recF; double* recT;      // by sid, 3 each
    unsigned long long* pairs; unsigned long long* pair_count; unsigned long long pair_cap;
    int* gmin; int* gmax;            // by sid, 3 each
};

}  // namespace demb200

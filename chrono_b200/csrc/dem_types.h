// =============================================================================
// dem_types.h -- shared host/device declarations of the B200 SMC-DEM engine.
//
// HBM layout (fp64; every per-sphere array is in STORAGE order = search-cell order of the last neighbour-list
// rebuild; two copies [0]/[1] that ping-pong every step, the live one is Ctrl::cur):
//   pos  : double4 (x, y, z, radius)                    32 B  one DRAM sector per neighbour gather
//   vel  : VelRec  (v xyz, omega xyz, sid, meta)        64 B  two aligned sectors, gathered only for real contacts
//   nl   : uint32[Kn][Np] column-major Verlet candidate list (storage slots of spheres within r_i+r_j+skin at the
//          last rebuild), each sphere's entries sorted by the partner's stable id so the summation order -- and
//          therefore every bit of the result -- does not depend on storage order, rebuild cadence or partition.
//          Mesh triangles within reach follow the sphere candidates in the same list (kTriFlag | triangle index,
//          ascending), so a sphere-triangle contact owns a history slot and a live bit exactly like a sphere pair.
//   hist : double4[Kn + nW][Np] column-major, ONE RECORD PER CANDIDATE SLOT (slot k of sphere s at k*Np + s, wall w
//          at (Kn+w)*Np + s): tangential displacement of that contact in canonical orientation (as stored by the
//          higher shape id, ChIterativeSolverMulticoreSMC.cpp:233-243) + steps in contact.  A 64-bit mask in the
//          sphere's VelRec says which slots are live.  No keys, no search: the address of a contact's history is
//          known as soon as the candidate is known, it is updated in place, and EVERY sphere keeps a record of EVERY
//          contact it takes part in (both partners hold bit-identical copies), so a thread only ever touches its own
//          column, records migrate with their sphere, and ghosts need no history.  When the lists are rebuilt the
//          live records travel through `stage` (keyed by partner shape id) into the slots of the new lists.
// sid = stable sphere id (= user index); shape id = shape_base + sid (Multicore numbering).
// =============================================================================
#pragma once
#include <cstddef>
#include <cstdint>

namespace demb200 {

constexpr int kMaxWalls = 16;
constexpr unsigned kEmptyKey = 0xFFFFFFFFu;
constexpr int kMaxSlots = 32;      // upper bound of K = simultaneous contacts of one sphere, walls included
constexpr int kMaxNeighbors = 64;  // upper bound of Verlet candidate slots Kn (live mask is 64 bits)
constexpr int kMaxMeshes = 16;     // triangle-mesh bodies (ChSystemDemMesh::AddMesh)
constexpr unsigned kTriFlag = 0x80000000u;  // candidate-list entry = triangle (index in the low bits), not a sphere slot
constexpr unsigned kHiFlag = 0x40000000u;   // sphere entry: the partner's stable id is higher than the owner's, i.e. the owner is body 1 of
                                            // the canonical (lower id, higher id) orientation -- decided once per list build so that the
                                            // force kernel does not have to gather the partner's id for it
// Tiled search grid (force kernel with shared-memory staging of the neighbour bins): search cells are numbered tile by tile,
// kTile x kTile x kTile cells per tile, x fastest inside a tile, so that the spheres of a tile are one contiguous run of the
// storage order and a tile plus its one-cell halo ("box", (kTile+2)^3 cells) holds every candidate of every sphere of the tile.
constexpr int kTile = 4;
constexpr int kTileCells = kTile * kTile * kTile;
constexpr int kBox = kTile + 2;
constexpr int kBoxCells = kBox * kBox * kBox;
constexpr unsigned kCodeHi = 0x8000u;     // nl16 entry: owner is body 1 (same meaning as kHiFlag)
constexpr unsigned kCodeRankBits = 6;     // nl16 entry = hi << 15 | box cell << 6 | rank in that cell; rank 63 = "not encodable"
constexpr unsigned kCodeNoRank = 63u;
constexpr int kCInfo = 13;                  // doubles per recorded contact (Buffers::cinfo)
constexpr unsigned kBndFlag = 0x80u;        // Buffers::ncnt bit 7: a candidate of this sphere is a ghost (slab mode) -> its forces wait for the halo
constexpr unsigned kEntryFlags = kHiFlag;  // flag bits of a sphere entry (everything that is not the storage slot)
constexpr unsigned kSlotHi = 0x80u;         // the same bit in the force kernel's shared-memory contact list (slot byte: slot < 64)

// device error bits (dem_b200 error codes are derived from these at sync points)
enum : unsigned {
    ERR_HISTORY_OVERFLOW = 4u,
    ERR_NEIGHBOR_OVERFLOW = 8u,
    ERR_NAN = 16u,
    ERR_PAIR_CAPACITY = 32u,
    ERR_MESH_CAPACITY = 64u,
    ERR_SKIN_EXCEEDED = 128u
};

enum WallType : int { WALL_BOX = 0, WALL_PLANE = 1, WALL_ZCYL = 2, WALL_SPHERE = 3, WALL_ZCONE = 4 };

// Composite material of a contact class, widened from the float values the reference computes
// (src/chrono/physics/ChContactMaterialSMC.cpp:107-130).
struct Comp {
    double E_eff, G_eff, mu, mu_roll, mu_spin, cr, adh, adh_dmt, adh_perko, kn, kt, gn, gt;
    // hoisted, material-only factors of the Hertz damping (ChIterativeSolverMulticoreSMC.cpp:283-287)
    double hertz_damp;  // -2*sqrt(5/6)*beta
    double gt_ratio;    // sqrt(St/Sn) = sqrt(4 G_eff / E_eff):  gt = gn * gt_ratio
};

struct Wall {
    int type;
    int enabled;       // DisableBCbyID / EnableBCbyID
    double pos[3];     // box centre / plane point / a point on the cylinder axis (world)
    double rot[4];     // box orientation (w,x,y,z)
    double hdims[3];   // box half dimensions / plane unit normal / (cylinder radius, +1 spheres inside | -1 outside, -) /
                       // (ball radius, +1 obstacle | -1 cavity, -) / cone (slope dz/dr, hmin, hmax), rot[0] = +1 spheres above
                       // the cone surface (inside a hopper) | -1 below it
    double vel[3];     // velocity of the wall body (moving boundaries)
    double omg[3];     // surface spin of a plane (Chrono::Dem SetBCPlaneRotation, ChSystemDem_impl.cpp:993-1001): the wall's material
    double rc[3];      // point at x moves with vel + omg x (x - rc); the geometry itself stays where it is
    double amin[3], amax[3];  // world AABB (ChCollisionSystemMulticore.cpp:395-406); boxes only
};

// All walls live in one device block so that they can move (SetBCOffsetFunction) without touching the CUDA graph.
struct WallSet {
    Wall w[kMaxWalls];
    double bb_min[3], bb_max[3];  // union of the box-wall AABBs (planes / cylinders are unbounded: excluded)
    int has_bb;
};

// One triangle-mesh body (Chrono::Dem "family"): rigid-body frame and velocity set by ApplyMeshMotion
// (src/chrono_dem/physics/ChSystemDem.cpp:1527-1561); triangles [tri_begin, tri_end) of the soup belong to it.
struct MeshBody {
    double pos[3], rot[4];   // body frame (w,x,y,z)
    double vel[3], omg[3];   // linear / angular velocity, world frame
    double mass;             // enters m_eff like any Multicore body mass
    unsigned tri_begin, tri_end;
};

// All meshes in one device block (moves without touching the CUDA graph, like WallSet).
struct MeshSet {
    MeshBody m[kMaxMeshes];
    unsigned long long bb[kMaxMeshes][6];  // order-preserving encoded world AABB of each mesh's triangles
    double wrench[kMaxMeshes][6];          // force and torque (about the body origin, world frame) of the last step
    int n;
    int enabled;                           // EnableMeshCollision
};

struct Params {
    unsigned N;   // spheres
    unsigned Np;  // N rounded up to a multiple of 32: pitch of the column-major arrays
    int nW;
    int K;        // most simultaneous contacts one sphere may have (staging slots at rebuild)
    int Kn;       // Verlet candidate slots per sphere
    int force_model, adhesion_model, tang_mode, use_mat_props, integrator;
    double char_vel, min_slip, min_roll, min_spin, dt;
    double g[3];
    double mass_coef, wall_mass;
    Comp comp[3];
    int bins[3];          // Multicore broadphase resolution (parity output only; the search grid is our own)
    unsigned shape_base;  // shape id of sphere sid is shape_base + sid
    double rmax;          // largest sphere radius
    double skin;          // Verlet skin: candidates are spheres closer than r_i + r_j + skin at rebuild time
    unsigned cell_cap;    // capacity of the search-cell arrays
    int track_wall_forces;  // accumulate the reaction force on every wall (GetBCReactionForces)
    int external_rebuild;   // slab mode: rebuilds happen only when the host asks (all ranks at the same step)
    int skin_adaptive;      // default skin only, not in slab mode: the skin grows (up to skin_max) while rebuilds come less than
    double skin_max;        // 12 steps apart (fast flow) and shrinks back to `skin` when they are more than 60 apart.  No
                            // result depends on the skin (summation order is by stable id), so this only moves cost around.
    double skin_tri;        // skin of the sphere-facet candidates (>= skin): a moving mesh (drum, mixer) sweeps much faster
                            // than the bed creeps, and only the spheres next to it pay for the longer facet lists
    int tiled;              // search cells numbered tile by tile + nl16 written: the tile force kernel can run
    unsigned nT;            // mesh triangles (shape ids nW .. nW + nT - 1; spheres follow: shape_base = nW + nT)
    unsigned tri_cap;       // capacity of the (search cell, triangle) pair list
};

constexpr unsigned FLAG_FIXED = 1u;
constexpr unsigned FLAG_GHOST = 2u;  // copy of a sphere owned by a neighbouring slab (multi-GPU): never integrated here

// 64-byte velocity record.  meta = flags (bits 0-7: 1 = fixed, 2 = ghost) | live wall-contact mask (bits 8-23);
// amask bit k = the contact with candidate k of the sphere's Verlet list carried force last step (history live).
struct __align__(16) VelRec {
    double v[3];
    double w[3];
    unsigned sid, meta;
    unsigned long long amask;
};

// Multicore broadphase grid of the current step (ChBroadphase.cpp:143-208)
struct GridDev {
    double origin[3], bin[3], inv[3];
    double wmin[kMaxWalls][3], wmax[kMaxWalls][3];  // wall AABBs offset by the grid origin (ChBroadphase.cpp:168-176)
};

// Device-resident control block: everything a step needs to know about itself, so the same CUDA graph can be
// replayed for every step (no host decision inside AdvanceSimulation).
struct Ctrl {
    unsigned cur;           // buffer index holding the state once all enqueued work has run
    unsigned rb_src;        // this step: buffer the rebuild kernels read (they write rb_src ^ 1)
    unsigned f_src;         // this step: buffer the force kernel reads (it writes f_src ^ 1)
    unsigned rebuild_now;   // this step rebuilds the search grid and the candidate lists
    unsigned need_rebuild;  // request (host: initialize / set_state)
    unsigned init_stage;    // first rebuild takes the contact history from Buffers::stage_init (checkpoint restart)
    unsigned err;           // sticky device error bits; alone in its 8-byte word (k_step_begin never overwrites it)
    unsigned err_pad_;
    unsigned long long nsteps, nrebuilds;
    unsigned long long bbox[6];   // order-preserving encoded doubles: min xyz, max xyz of the sphere AABBs
    unsigned long long sbox[6];   // same encoding: AABB of the spheres alone (local ones and ghosts) at the last list rebuild -- the
    unsigned long long sbox_next[6];  // search grid covers this box (+ the meshes), not the walls: a slab of a long box must not
    unsigned sbox_pending, sbox_pad_; // bin the whole box.  sbox_next is gathered by k_bin_count while it reads every position anyway.
    unsigned long long max_dx2;   // raw bits of max |x_new - x_old|^2 over the spheres, last step
    unsigned long long import_dx2;  // same, for positions replaced by the host since the last step (dem_b200_set_state): uses up skin like a step
    double travel;                // sum of per-step max displacements since the last rebuild
    double last_dx;               // max displacement of the step before the last one (growth estimate of the slab vote)
    double skin;                  // Verlet skin the current lists were built with (Params::skin unless adaptive)
    unsigned since_rebuild;       // steps since the lists were built
    unsigned max_cand;            // longest candidate list of the last rebuild (the adaptive skin backs off near Kn)
    unsigned n_bnd;               // slab mode: owned spheres with a ghost among their candidates (Buffers::bnd_list), last rebuild
    unsigned n_bnd_pad_;
    double travel_mesh;           // how far mesh vertices moved since the last rebuild (ApplyMeshMotion); counts against the
                                  // facet candidates' own, larger skin (Params::skin_tri) on top of `travel`
    // search grid (cells >= 2 rmax + skin), x fastest
    double s_org[3], s_inv[3];
    int s_dim[3];
    int s_perm[3];          // axis order of the cell numbering: s_perm[0] runs fastest.  (0,1,2) on one GPU; slab mode puts the slab axis x
                            // last, so that ghosts and ghost senders (the layers next to the slab faces) are contiguous runs of the storage order
    int t_dim[3];           // tiles per axis (Params::tiled): ceil(s_dim / kTile); s_ncell then counts the padded cells
    unsigned s_ncell;
    GridDev mc;
    unsigned long long n_contacts, pair_count;
    double wall_force[kMaxWalls][3];  // force exerted by the spheres on wall w during the last step
};

// Slab decomposition (one process per GPU): counters and index lists of the neighbour exchange.
struct SlabDev {
    unsigned n_keep;        // owned spheres that stay (extract)
    unsigned n_out[2];      // spheres leaving to the left / right neighbour
    unsigned n_gsend[2];    // owned spheres whose copies the left / right neighbour needs as ghosts
    unsigned want_rebuild;  // this rank's Verlet skin is (about to be) used up
    // device-driven rebuild (P2P mode): what arrived, and the resulting layout [owned | ghosts from left | from right]
    unsigned n_in[2];       // migrants received from the left / right neighbour
    unsigned g_in[2];       // ghosts received from the left / right neighbour
    unsigned n_own;         // n_keep + n_in[0] + n_in[1]
};

// Direct peer-to-peer halo (slab mode, one process per GPU on one NVSwitch box).  Every rank exposes one region through
// CUDA IPC: this control block followed by the halo landing buffers [from left | from right] x [step parity].  A rank's
// pack kernel STORES its boundary spheres' state straight into the neighbour's landing buffer over NVLink and then
// publishes the step number in `arrive`; the neighbour's unpack kernel waits for that number.  The per-step "rebuild
// now?" vote is a store of (step << 1 | flag) into slot [step & 7][rank] of every peer; no collective, no host.
constexpr int kMaxRanks = 8;
struct P2PCtl {
    unsigned long long arrive[2][2];         // [0: from the left neighbour, 1: from the right][parity] = step that landed
    unsigned long long vote[8][kMaxRanks];   // [step & 7][rank] = (step << 1) | wants_rebuild
    // rebuild: migrating spheres and the new ghost set are stored into the neighbour's landing buffers as well
    unsigned long long mig_arrive[2], gho_arrive[2];  // [from side] = number of the rebuild whose records landed
    unsigned mig_count[2], gho_count[2];              // how many
};
constexpr size_t kP2PCtlBytes = 1024;        // control block padded; landing buffers follow

struct P2PDev {
    P2PCtl* self;                  // my own control block
    double* land[2][2];            // my landing buffers [from side][parity]
    P2PCtl* peer[kMaxRanks];       // every rank's control block as mapped here (peer[rank] == self)
    double* peer_land[2][2];       // [0: left neighbour, 1: right neighbour][parity]: THEIR buffer for data coming from me
    unsigned* done;                // block counters of the pack kernels (2)
    unsigned long long* host_vote; // pinned, mapped: [step & 7] = (step << 1) | any rank wants a rebuild
    int rank, world, ahead;
    unsigned long long first_step; // first step that ran in P2P mode: votes of earlier steps do not exist
    double* mig_land[2];           // my landing buffers for migrants / ghost records [from side]
    double* gho_land[2];
    double* peer_mig[2];           // [0: left neighbour, 1: right]: THEIR landing buffer for records coming from me
    double* peer_gho[2];
    unsigned cap_mig, cap_gho;     // records per landing buffer
};

struct Buffers {
    Ctrl* ctrl;
    WallSet* walls;
    SlabDev* slab;
    uint32_t* send_pre[2];   // pre-sort index of every ghost-sender, in message order
    uint32_t* send_slot[2];  // its storage slot after the sort (per-step pack list)
    uint32_t* ghost_slot[2]; // storage slot of the i-th ghost received from the left / right neighbour
    uint32_t* inv_perm;      // scratch: pre-sort index -> storage slot
    uint32_t* bnd_list;      // storage slots of the spheres whose candidate list holds a ghost (any order), Ctrl::n_bnd of them
    // state, ping-pong
    double4* pos[2];
    VelRec* vel[2];
    double* acc[2];       // previous-step acceleration (Chung only), 6 per sphere
    double4* hist;        // [Kn + nW][Np], updated in place
    double* hrel;         // [Kn + nW][Np] initial normal speed (only Hooke/Flores with material properties)
    // history in transit during a rebuild: (disp xyz, partner shape id | steps << 32), keyed, compact
    double4* stage; double* stage_rel; uint32_t* stage_cnt;                  // [K][Np], [K][Np], [Np]
    double4* stage_init; double* stage_rel_init; uint32_t* stage_cnt_init;   // same, user order; null unless add_history
    // neighbour search
    uint32_t* cell; uint32_t* rank; uint32_t* perm;
    uint32_t* cell_count; uint32_t* cell_start; uint32_t* block_sums;
    uint32_t* nl;         // [Kn][Np]
    uint16_t* nl16;       // [Kn][Np] (Params::tiled) the sphere candidates again, as (box cell, rank in cell) relative to the owner's tile
    uint32_t* ncnt;       // [Np]
    // recording (parity tests / smoke)
    double* recF; double* recT;      // by sid, 3 each
    unsigned long long* pairs; unsigned long long pair_cap;
    int* gmin; int* gmax;            // by sid, 3 each: Multicore HashMin / HashMax of the sphere AABB
    // per-contact records (Chrono::Dem SetRecordingContactInfo, ChSystemDem_impl.cpp:443-635): kCInfo doubles per history
    // slot, [(slot * Np + s) * kCInfo + c] = force on the sphere: normal part xyz, tangential part xyz; rolling + spinning
    // resistance torque xyz; v_rot xyz; characteristic collision time.  Null unless dem_b200_enable_contact_info.
    double* cinfo;
    // triangle meshes: soup in the body frame and in the world frame (9 doubles per triangle: A, B, C), owner mesh,
    // and the CSR "triangles reaching search cell c" of the last rebuild
    MeshSet* meshes;
    double* tri_loc; double* tri_w; uint32_t* tri_mesh;
    uint32_t* tcell_count; uint32_t* tcell_start; uint32_t* tblock_sums; uint32_t* tcell_tri;
};

}  // namespace demb200

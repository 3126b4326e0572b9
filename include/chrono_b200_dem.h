/* =============================================================================
 * chrono_b200_dem.h -- C ABI of the B200-native smooth-contact (SMC) granular DEM engine.
 *
 * This is the drop-in boundary for the hot path of Chrono::Dem (a.k.a. chrono_gpu):
 *     chrono::dem::ChSystemDem::Initialize / AdvanceSimulation / getters
 *     (reference: src/chrono_dem/physics/ChSystemDem.h:40-393; impl loop src/chrono_dem/gpu/ChDemSMC.cu:619-691)
 * The reference has no C ABI of its own (its seam is the pimpl pointer ChSystemDem::m_sys, ChSystemDem.h:350);
 * these entry points are what a maintainer binds in place of ChSystemDem_impl -- see INTEGRATION.md.  The C++
 * mirror of the reference class (include/chrono_dem/physics/ChSystemDem.h) is written on top of exactly this ABI.
 *
 * Arithmetic contract: fp64 state, the force law and collision semantics of Chrono::Multicore's
 * ChSystemMulticoreSMC (src/chrono_multicore/solver/ChIterativeSolverMulticoreSMC.cpp:56-546,
 * src/chrono/collision/multicore/*), which is the parity oracle (oracle/).
 *
 * All pointers are HOST pointers unless the name ends in _dev.  Every function returns 0 on success or a
 * negative DEMB200_E* code; dem_b200_last_error() gives the text.  No function calls exit().  No CPU fallback:
 * if no CUDA device / kernel image is available every entry point fails with DEMB200_ECUDA.
 * ============================================================================= */
#ifndef CHRONO_B200_DEM_H
#define CHRONO_B200_DEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dem_b200_system dem_b200_system; /* opaque */

enum {
    DEMB200_OK = 0,
    DEMB200_ECUDA = -1,    /* CUDA runtime error (no device, launch failure, ...) */
    DEMB200_EINVAL = -2,   /* bad argument / call order */
    DEMB200_EGRID = -3,    /* reserved (the search grid adapts itself; kept for ABI stability) */
    DEMB200_EHISTORY = -4, /* a sphere has more simultaneous contacts than history_slots */
    DEMB200_ENAN = -5,     /* non-finite state detected */
    DEMB200_ECAPACITY = -6, /* a recording buffer (pairs) or the mesh cell list overflowed */
    DEMB200_ENEIGHBORS = -7 /* a sphere has more neighbour candidates than neighbor_slots */
};

/* ChSystemSMC::ContactForceModel / AdhesionForceModel / TangentialDisplacementModel
 * (src/chrono/physics/ChSystemSMC.h:34-53); CHDEM_FRICTION_MODE FRICTIONLESS/SINGLE_STEP/MULTI_STEP map to
 * DEMB200_TANG_NONE/ONESTEP/MULTISTEP (src/chrono_dem/ChDemDefines.h:47). */
enum { DEMB200_HOOKE = 0, DEMB200_HERTZ = 1, DEMB200_PLAINCOULOMB = 2, DEMB200_FLORES = 3 };
enum { DEMB200_ADH_CONSTANT = 0, DEMB200_ADH_DMT = 1, DEMB200_ADH_PERKO = 2 };
enum { DEMB200_TANG_NONE = 0, DEMB200_TANG_ONESTEP = 1, DEMB200_TANG_MULTISTEP = 2 };
/* CHDEM_TIME_INTEGRATOR, same order as src/chrono_dem/ChDemDefines.h:44.  CENTERED_DIFFERENCE is the
 * semi-implicit Euler of Chrono::Multicore (src/chrono/physics/ChBody.cpp:288-310) and the parity integrator. */
enum { DEMB200_FORWARD_EULER = 0, DEMB200_CHUNG = 1, DEMB200_CENTERED_DIFFERENCE = 2, DEMB200_EXTENDED_TAYLOR = 3 };

/* ChContactMaterialSMC (src/chrono/physics/ChContactMaterialSMC.h:85-98); float, as in the reference. */
typedef struct dem_b200_material {
    float young, poisson;
    float mu_s, mu_roll, mu_spin, cr;
    float adhesion, adhesion_dmt, adhesion_perko;
    float kn, kt, gn, gt;
} dem_b200_material;

enum { DEMB200_MAT_SPHERE = 0, DEMB200_MAT_WALL = 1, DEMB200_MAT_MESH = 2 };

typedef struct dem_b200_config {
    int device;           /* CUDA device ordinal */
    int force_model, adhesion_model, tangential_mode, use_mat_props;
    int integrator;
    int history_slots;    /* contact records per sphere, walls included (reference Dem: 12, Multicore: 20 on the
                             higher-id body only); every sphere keeps a record of each of its contacts; 0 -> 16, max 32 */
    double char_vel, min_slip_vel, min_roll_vel, min_spin_vel; /* ChSettings.h:121-124 */
    double dt;
    double gravity[3];
    int bins_per_axis[3]; /* Multicore broadphase resolution (collision_settings::bins_per_axis): defines the bin ids
                             reported by dem_b200_get_bins/get_grid; the engine's own search grid is independent */
    dem_b200_material material[3]; /* [DEMB200_MAT_SPHERE|WALL|MESH] */
    double mass_coef;     /* sphere mass = mass_coef * (r*r*r)  (= 4/3 pi rho for solid spheres) */
    double wall_mass;     /* mass of the body carrying the walls; enters m_eff (Multicore Q9) */
    double mesh_mass;     /* default mass of mesh bodies */
    double verlet_skin;   /* neighbour candidates = spheres closer than r_i + r_j + skin when the lists are built; the
                             lists are rebuilt (on the device, no host round trip) before any sphere can have moved
                             skin/2.  < 0 -> default: 0.25 * largest radius, adapting up to 0.5 on the device while rebuilds come less than 12
                             steps apart (0.35, fixed, in slab mode); 0 -> rebuild every step */
    int neighbor_slots;   /* candidate slots per sphere; 0 -> 32, max 64 */
} dem_b200_config;

/* ---- lifecycle ---------------------------------------------------------------------------------------- */
int dem_b200_create(const dem_b200_config* cfg, dem_b200_system** out);
void dem_b200_destroy(dem_b200_system* s);
const char* dem_b200_last_error(const dem_b200_system* s); /* s may be NULL: last create() error */
int dem_b200_set_config(dem_b200_system* s, const dem_b200_config* cfg); /* model/material/dt changes mid-run */

/* ---- scene (before initialize) -- ChSystemDem::SetParticles / CreateBC* ---------------------------------- */
/* n spheres in user order; vel3/omega3/fixed may be NULL (zeros).  radius has n entries. */
int dem_b200_set_spheres(dem_b200_system* s, size_t n, const double* pos3, const double* vel3,
                         const double* omega3, const double* radius, const uint8_t* fixed);
/* Box wall (Multicore box shape on the fixed wall body): world centre, quaternion (w,x,y,z), half dims.
 * Returns the wall index (>= 0) or an error (< 0).  Shape id of wall k is k; spheres follow (Q12 numbering). */
int dem_b200_add_box_wall(dem_b200_system* s, const double pos[3], const double rot[4], const double hdims[3]);
/* Infinite plane wall (Chrono::Dem CreateBCPlane, src/chrono_dem/physics/ChSystemDem.h:228): point + unit normal
 * pointing into the domain.  Contact iff distance < r; eff. radius = r. */
int dem_b200_add_plane_wall(dem_b200_system* s, const double pos[3], const double normal[3]);
/* Z-axis cylinder (Chrono::Dem CreateBCCylinderZ, src/chrono_dem/physics/ChSystemDem.h:235): spheres_inside != 0 ->
 * container wall (normal towards the axis), else an obstacle. */
int dem_b200_add_zcylinder_wall(dem_b200_system* s, const double center[3], double radius, int spheres_inside);
/* Ball (Chrono::Dem CreateBCSphere, src/chrono_dem/physics/ChSystemDem.h:206): spheres_outside != 0 -> obstacle, identical
 * to Multicore's sphere_sphere against a fixed sphere of that radius; else a spherical cavity holding the spheres. */
int dem_b200_add_sphere_wall(dem_b200_system* s, const double center[3], double radius, int spheres_outside);
/* Cone about the z axis (Chrono::Dem CreateBCConeZ, src/chrono_dem/physics/ChSystemDem.h:212): surface
 * z - tip.z = slope * rho, active for hmin < z < hmax (world z); spheres_above != 0 -> hopper (spheres above the surface). */
int dem_b200_add_zcone_wall(dem_b200_system* s, const double tip[3], double slope, double hmin, double hmax, int spheres_above);
/* Move a wall (SetBCOffsetFunction, src/chrono_dem/physics/ChSystemDem_impl.cpp:863-948): new reference point and
 * velocity (either may be NULL).  Legal between steps; does not invalidate the step graph. */
int dem_b200_set_wall_state(dem_b200_system* s, int wall, const double pos[3], const double vel[3]);
int dem_b200_set_wall_velocity(dem_b200_system* s, int wall, const double pos[3], const double vel[3]); /* alias */
int dem_b200_enable_wall(dem_b200_system* s, int wall, int enabled); /* DisableBCbyID / EnableBCbyID */
/* Reaction force on a wall during the last step (GetBCReactionForces, ChSystemDem_impl.cpp:1003-1029). */
int dem_b200_track_wall_forces(dem_b200_system* s, int enable);
int dem_b200_wall_force(dem_b200_system* s, int wall, double force[3]);

/* Explicit coefficients of a contact class (DEMB200_MAT_SPHERE = sphere-sphere, _WALL = sphere-wall, _MESH =
 * sphere-mesh), overriding the composite of the two ChContactMaterialSMC materials.  This is how the per-class setters
 * of Chrono::Dem map (SetKn_SPH2WALL, SetStaticFrictionCoeff_SPH2MESH, ... src/chrono_dem/physics/ChSystemDem.h:107-160).
 * c == NULL removes the override. */
typedef struct dem_b200_contact_class {
    double E_eff, G_eff;          /* used when use_mat_props */
    double mu, mu_roll, mu_spin, cr;
    double adhesion;              /* constant adhesion force */
    double kn, kt, gn, gt;        /* used when !use_mat_props (Multicore convention) */
} dem_b200_contact_class;
int dem_b200_set_contact_class(dem_b200_system* s, int cls, const dem_b200_contact_class* c);

/* ---- triangle meshes -- ChSystemDemMesh (src/chrono_dem/physics/ChSystemDem.h:398-523) ------------------------
 * A mesh is a rigid body carrying one Multicore triangle shape per facet (ChNarrowphasePRIMS.cpp:379-437
 * triangle_sphere: one-sided, the outside is where (B-A)x(C-A) points).  Shape ids: walls 0..nW-1, then the triangles
 * of all meshes in insertion order, then the spheres.  All meshes must be added before initialize. */
/* AddMesh (ChSystemDem.h:413): ntri triangles, 9 doubles each (A, B, C in the mesh frame); mass enters m_eff
 * (<= 0 -> cfg.mesh_mass).  Returns the mesh id (>= 0) or an error. */
int dem_b200_add_mesh(dem_b200_system* s, size_t ntri, const double* verts9, double mass);
int dem_b200_num_meshes(const dem_b200_system* s);
size_t dem_b200_num_triangles(const dem_b200_system* s);
/* ApplyMeshMotion (ChSystemDem.h:436, ChSystemDem.cpp:627-630): frame position, orientation (w,x,y,z), linear and
 * angular velocity in the world frame.  Any pointer may be NULL (unchanged).  Legal between steps; asynchronous. */
int dem_b200_set_mesh_motion(dem_b200_system* s, int mesh, const double pos[3], const double rot[4],
                             const double lin_vel[3], const double ang_vel[3]);
int dem_b200_enable_mesh_collision(dem_b200_system* s, int enabled); /* EnableMeshCollision (ChSystemDem.h:430) */
/* CollectMeshContactForces (ChSystemDem.h:493-496, ChSystemDem.cpp:1752-1766): force on the mesh and torque about its
 * frame origin (world frame) exerted by the spheres during the last step. */
int dem_b200_mesh_wrench(dem_b200_system* s, int mesh, double force[3], double torque[3]);

/* ---- run -- ChSystemDem::Initialize / AdvanceSimulation --------------------------------------------------- */
int dem_b200_initialize(dem_b200_system* s);
/* nsteps steps of size dt, asynchronous on the system's stream; errors surface at the next sync point. */
int dem_b200_step(dem_b200_system* s, int nsteps);
int dem_b200_sync(dem_b200_system* s);
/* Same, bracketed by CUDA events on the launching stream; *ms = device time of the nsteps steps. */
int dem_b200_step_timed(dem_b200_system* s, int nsteps, float* ms);
/* Per-kernel device time (ms, summed over nsteps) in the order of dem_b200_kernel_name(); n_out <= 16. */
int dem_b200_step_profile(dem_b200_system* s, int nsteps, float* ms_per_kernel, int* n_out);
const char* dem_b200_kernel_name(int k);
/* Host-buffer round trip of the hot path: upload state (user order), nsteps steps, download state.
 * This is the call timed as "e2e" by bench.py.  Any of the out pointers may be NULL. */
int dem_b200_advance_host(dem_b200_system* s, size_t n, const double* pos3_in, const double* vel3_in,
                          const double* omega3_in, int nsteps, double* pos3_out, double* vel3_out,
                          double* omega3_out);

/* ---- state access (user order) -- GetParticlePosition / Velocity / AngVelocity ------------------------------- */
size_t dem_b200_num_spheres(const dem_b200_system* s);
int dem_b200_get_state(dem_b200_system* s, double* pos3, double* vel3, double* omega3);
/* New positions use up Verlet skin like a time step does: the largest jump of a sphere is added to the travel since the last
 * neighbour-list rebuild, so handing back positions that did not move (a co-simulation round trip) costs no rebuild. */
int dem_b200_set_state(dem_b200_system* s, const double* pos3, const double* vel3, const double* omega3);
/* The next step rebuilds the neighbour lists whatever the travel (no result depends on it; used by profiling scripts). */
int dem_b200_request_rebuild(dem_b200_system* s);
int dem_b200_get_sphere(dem_b200_system* s, size_t i, double pos[3], double vel[3], double omega[3]);
/* Linear acceleration (gravity included) each sphere had in the last step, user order -- GetParticleLinAcc
 * (ChSystemDem.h:273, ChSystemDem_impl.cpp:1290-1296) and the fx,fy,fz columns of WriteParticleFile (:322-327).  Zero
 * before the first step and right after dem_b200_set_state; zero for fixed spheres.  Not available in slab mode. */
int dem_b200_get_accel(dem_b200_system* s, double* acc3);
int dem_b200_get_sphere_accel(dem_b200_system* s, size_t i, double acc[3]);
double dem_b200_time(const dem_b200_system* s);

/* ---- reductions -- GetMaxParticleZ, GetParticlesKineticEnergy, ... (ChSystemDem.h:246-262) ------------------ */
enum { DEMB200_RED_MAX_Z = 0, DEMB200_RED_MIN_Z = 1, DEMB200_RED_KE = 2, DEMB200_RED_MAX_SPEED = 3,
       DEMB200_RED_COUNT_ABOVE_Z = 4, DEMB200_RED_COUNT_ABOVE_X = 5,
       DEMB200_RED_NUM_CONTACTS = 6, /* sum over the owned spheres of their force-carrying contacts (MultiStep only) */
       DEMB200_RED_KE_TRANSLATIONAL = 7 /* sum of m v^2 / 2 only: what Chrono::Dem's GetParticlesKineticEnergy returns
                                           (ChSystemDem_impl.cpp:1250-1264); DEMB200_RED_KE adds the rotational part */ };
int dem_b200_reduce(dem_b200_system* s, int which, double arg, double* out);

/* ---- parity / inspection (tests, smoke) ---------------------------------------------------------------------- */
/* When enabled, each step also records per-sphere contact force & torque, the contact-pair list and the
 * per-sphere bin ranges of that step.  max_pairs = capacity of the pair buffer. */
int dem_b200_enable_recording(dem_b200_system* s, int enable, size_t max_pairs);
int dem_b200_get_forces(dem_b200_system* s, double* force3, double* torque3); /* user order */
/* pair key = (shapeA << 32) | shapeB with shapeA < shapeB; wall k is shape k, sphere i is shape num_walls + i */
int dem_b200_get_pairs(dem_b200_system* s, uint64_t* pairs, size_t capacity, size_t* n);
int dem_b200_get_bins(dem_b200_system* s, int32_t* gmin3, int32_t* gmax3); /* user order */
int dem_b200_get_grid(dem_b200_system* s, double origin[3], double bin_size[3], double inv_bin_size[3]);
/* history rows: owner shape id, other shape id, displacement, duration, initial normal speed */
int dem_b200_get_history(dem_b200_system* s, uint32_t* owner, uint32_t* other, double* disp3, double* duration,
                         double* relvel_init, size_t capacity, size_t* n);
int dem_b200_add_history(dem_b200_system* s, uint32_t owner_shape, uint32_t other_shape, const double disp[3],
                         double duration, double relvel_init);
int dem_b200_num_walls(const dem_b200_system* s);
/* steps executed, neighbour-list rebuilds among them, contacts (sphere-sphere counted from both sides + sphere-wall)
 * seen by the last recorded step */
int dem_b200_get_stats(dem_b200_system* s, unsigned long long* nsteps, unsigned long long* nrebuilds,
                       unsigned long long* contacts_last_step);

/* ---- slab decomposition: one process per GPU (no reference counterpart: Chrono::Dem is single-GPU and the
 * MPI module Chrono::Distributed was removed, CHANGELOG.md:496-501) --------------------------------------------------
 * The engine keeps owned spheres and ghosts (copies of the neighbour slabs' spheres within 2 r_max + skin of the
 * face) in one array; ghosts are never integrated.  Between neighbour-list rebuilds the halo is a fixed-index
 * pack -> send/recv -> unpack of (pos, v, omega); spheres change owner only at a rebuild, together with their contact
 * history.  No force or history exchange: both partners of a contact keep their own copy.  The caller owns the device
 * buffers and the transport (chrono_b200/slab.py: torch.distributed, NCCL).  Call order of a rebuild:
 *   extract -> [exchange migrants] -> append(ghost=0) x2 -> select_ghosts -> [exchange ghosts] -> append(ghost=1, dir 0)
 *   -> append(ghost=1, dir 1) -> finish_rebuild -> step ... ; per step: pack x2 -> [exchange] -> unpack x2 -> step. */
int dem_b200_set_stream(dem_b200_system* s, void* cuda_stream);              /* before initialize: run on the caller's stream */
int dem_b200_set_sphere_ids(dem_b200_system* s, const uint32_t* ids);       /* global stable ids of the spheres of set_spheres */
int dem_b200_mgpu_enable(dem_b200_system* s, size_t capacity, double rmax_global); /* before initialize */
int dem_b200_mgpu_sizes(dem_b200_system* s, size_t* halo_bytes, size_t* ghost_bytes, size_t* migrant_bytes, double* cut);
int dem_b200_mgpu_extract(dem_b200_system* s, double lo, double hi, void* out_left_dev, void* out_right_dev,
                          size_t cap_records, size_t* n_keep, size_t* n_left, size_t* n_right);
int dem_b200_mgpu_append(dem_b200_system* s, const void* in_dev, size_t n, int ghost, int dir);
int dem_b200_mgpu_select_ghosts(dem_b200_system* s, double lo, double hi, double cut, void* out_left_dev, void* out_right_dev,
                                size_t cap_records, size_t* n_left, size_t* n_right);
int dem_b200_mgpu_finish_rebuild(dem_b200_system* s);
int dem_b200_mgpu_counts(dem_b200_system* s, size_t* n_own, size_t* n_ghost_left, size_t* n_ghost_right, size_t* n_send_left,
                         size_t* n_send_right);
int dem_b200_mgpu_pack(dem_b200_system* s, int dir, void* out_dev);
int dem_b200_mgpu_unpack(dem_b200_system* s, int dir, const void* in_dev);
int dem_b200_mgpu_want_rebuild(dem_b200_system* s, int* flag_dev); /* writes 1/0 to DEVICE memory (for an all-reduce) */
/* Same for a driver that acts on the answer `steps_ahead` steps late (no host sync per step): those steps are assumed to
 * move the spheres as far as the last one did; if they move further the next step fails loudly (DEMB200_EINVAL). */
int dem_b200_mgpu_want_rebuild_ahead(dem_b200_system* s, int* flag_dev, int steps_ahead);
/* Direct peer-to-peer halo for ranks on one NVLink/NVSwitch box (replaces pack -> send/recv -> unpack and the per-step
 * all-reduce): every rank exports one region (control block + landing buffers for max_records ghosts per side) as a
 * 64-byte CUDA IPC handle; after the handles of all ranks were gathered (any transport), import() maps them.  From then on
 * dem_b200_step() itself moves the halo: the pack kernel stores into the neighbour's landing buffer and publishes the step
 * number, the unpack kernel waits for it; the rebuild vote of every step is stored to all peers and OR-ed on the device.
 * The host only polls the vote (steps_ahead >= 1 steps late, so that the GPUs never wait for it) and runs the rebuild
 * protocol above when it says so. */
int dem_b200_p2p_export(dem_b200_system* s, size_t max_records, void* handle64);
int dem_b200_p2p_import(dem_b200_system* s, int rank, int world, const void* handles /* world x 64 bytes */, int steps_ahead);
int dem_b200_p2p_poll_vote(dem_b200_system* s, unsigned long long step, int* flag);
/* The rebuild protocol above (extract .. finish_rebuild) without a collective or a per-message host round trip: migrants
 * and ghost records are stored into the neighbours' landing buffers, counts and arrival flags follow; one host
 * synchronisation at the end.  Every rank calls it at the same step.  counts (may be NULL) = owned spheres, ghosts
 * received from left / right, owned spheres sent as ghosts to left / right, spheres migrated out to left / right. */
int dem_b200_p2p_rebuild(dem_b200_system* s, double lo, double hi, size_t counts[7]);
unsigned long long dem_b200_step_count(const dem_b200_system* s); /* time steps enqueued so far (numbering of the votes) */

/* owned spheres (ghosts excluded) in arbitrary order: global id, pos, vel, omega to HOST buffers */
int dem_b200_export_owned(dem_b200_system* s, uint32_t* sid, double* pos3, double* vel3, double* omega3, size_t capacity,
                          size_t* n);

/* The inverse, for the host round trip of a slab engine (bench.py e2e at N > 1): the n records of the LAST
 * dem_b200_export_owned, same order, values possibly changed by the caller, go back into their spheres.  Must directly
 * follow that export (no step, no rebuild in between).  With new positions the candidate lists and the neighbours' ghost
 * copies are stale: run the slab rebuild protocol before the next step. */
int dem_b200_import_owned(dem_b200_system* s, size_t n, const double* pos3, const double* vel3, const double* omega3);
/* ---- per-contact records -- SetRecordingContactInfo / getNormalForce ... / WriteContactInfoFile
 * (src/chrono_dem/physics/ChSystemDem.h:180,326-338; ChSystemDem_impl.cpp:443-635).  While enabled every step also keeps,
 * for each force-carrying contact of each sphere, 13 doubles: the force on that sphere split into its normal (xyz) and
 * tangential (xyz) part, the rolling + spinning resistance torque on it (xyz), v_rot (xyz) and the characteristic collision
 * time.  Needs MultiStep friction (the contact map); steps run un-graphed through the recording kernels. */
int dem_b200_enable_contact_info(dem_b200_system* s, int enable);
/* One contact of sphere `sphere` (user index): other_shape = shape id of the partner (wall k -> k, facet t -> num_walls + t,
 * sphere j -> num_walls + num_triangles + j).  *found = 0 and zeros when the two do not touch. */
int dem_b200_get_contact_info(dem_b200_system* s, size_t sphere, uint32_t other_shape, double info13[13], int* found);
/* All sphere-sphere contacts, each once (bi < bj, the record is the one of sphere bi), any order.  Pass NULL arrays to get
 * the count. */
int dem_b200_get_contact_infos(dem_b200_system* s, uint32_t* bi, uint32_t* bj, double* info13, size_t capacity, size_t* n);
/* Surface spin of a wall -- SetBCPlaneRotation (ChSystemDem.h:238, ChSystemDem_impl.cpp:993-1001): the wall's material point
 * at a contact moves with vel + omega x (x - center); the geometry does not move.  Legal at any time. */
int dem_b200_set_wall_rotation(dem_b200_system* s, int wall, const double center[3], const double omega[3]);

/* Device error bits (history / neighbour overflow, skin exceeded, NaN ...) are sticky: every later sync point reports them.
 * Once the cause is dealt with (set_config with more slots, set_state ...) this clears them. */
int dem_b200_clear_error(dem_b200_system* s);

#ifdef __cplusplus
}
#endif
#endif /* CHRONO_B200_DEM_H */

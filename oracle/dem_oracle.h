/* =============================================================================
 * oracle/dem_oracle.h -- TEST INFRASTRUCTURE, not part of the product path.
 *
 * C interface of the CPU oracle: an fp64 restatement of the Chrono::Multicore smooth-contact
 * (SMC) time step  ChSystemMulticoreSMC::DoStepDynamics  for bodies carrying sphere, box and
 * triangle collision shapes.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libchrono_b200_dem.so) never does.
 *
 * Parity status: PINNED for sphere_sphere / box_sphere / snap_to_box (reference golden vectors,
 * src/tests/unit_tests/collision/utest_COLL_narrow_prims.cpp:39-73,974-1226) and, in the build
 * container, bit-for-bit against the reference's own objects compiled into oracle/_ref
 * (broadphase, narrowphase, function_CalcContactForces).  triangle_sphere has no golden vector
 * in the reference; it is pinned only against oracle/_ref.
 * ============================================================================= */
#ifndef DEM_ORACLE_H
#define DEM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ChSystemSMC enums, src/chrono/physics/ChSystemSMC.h:34-53 (same numeric values) */
enum { ORC_HOOKE = 0, ORC_HERTZ = 1, ORC_PLAINCOULOMB = 2, ORC_FLORES = 3 };
enum { ORC_ADH_CONSTANT = 0, ORC_ADH_DMT = 1, ORC_ADH_PERKO = 2 };
enum { ORC_TANG_NONE = 0, ORC_TANG_ONESTEP = 1, ORC_TANG_MULTISTEP = 2 };

/* ChContactMaterialSMC, src/chrono/physics/ChContactMaterialSMC.h:85-98 + ChContactMaterial.h:94-98
 * (all members are float in the reference) */
typedef struct OrcMaterial {
    float young, poisson;
    float mu_s, mu_roll, mu_spin, cr;
    float adhesion, adhesion_dmt, adhesion_perko;
    float kn, kt, gn, gt;
} OrcMaterial;

/* solver_settings / collision_settings subset, src/chrono_multicore/ChSettings.h:41-48,117-124 */
typedef struct OrcSettings {
    int force_model, adhesion_model, tangential_mode, use_mat_props;
    double char_vel, min_slip_vel, min_roll_vel, min_spin_vel;
    double dt;
    double gravity[3];
    int bins_per_axis[3];
    int num_threads; /* 0 = omp default */
} OrcSettings;

void* orc_create(const OrcSettings* s);
void orc_destroy(void* h);
void orc_set_settings(void* h, const OrcSettings* s);

int orc_add_material(void* h, const OrcMaterial* m);
int orc_add_body(void* h, double mass, const double inertia[3], const double pos[3], const double rot[4],
                 const double vel[3], const double omega_loc[3], int fixed);
/* n bodies, one sphere shape each at the body origin; inertia = 2/5 m r^2.  Returns first body id. */
int orc_add_spheres(void* h, int n, const double* pos3, const double* vel3, const double* omega3,
                    const double* radius, const double* mass, int material);
int orc_add_box(void* h, int body, int material, const double lpos[3], const double lrot[4], const double hdims[3]);
int orc_add_triangles(void* h, int body, int material, int n, const double* verts9);
void orc_set_body_state(void* h, int body, const double pos[3], const double rot[4], const double vel[3],
                        const double omega_loc[3]);
void orc_set_body_fixed(void* h, int body, int fixed);

/* Advance nsteps steps.  Returns 0, or 1 if a body ran out of the 20 history slots (max_shear). */
int orc_step(void* h, int nsteps);
/* Collision detection + force evaluation only (no state update); fills contacts/forces for queries. */
int orc_eval(void* h);

int orc_num_bodies(void* h);
int orc_num_shapes(void* h);
void orc_get_body_state(void* h, double* pos3, double* rot4, double* vel3, double* omega3);
/* world-frame shape AABBs as GenerateAABB produces them (before the broadphase offsets them) */
void orc_generate_aabb(void* h, double* min3, double* max3);
void orc_get_grid(void* h, double origin[3], double bin_size[3], double inv_bin_size[3], int bins[3]);
void orc_get_shape_bins(void* h, int* gmin3, int* gmax3);
/* sizes[0]=num_active_bins sizes[1]=num_bin_aabb_intersections sizes[2]=num_possible_collisions */
void orc_get_broadphase_sizes(void* h, long long sizes[3]);
void orc_get_bin_csr(void* h, unsigned* bin_active, unsigned* bin_start_index, unsigned* bin_aabb_number);
void orc_get_pairs(void* h, long long* pair_shape_ids);
long long orc_num_contacts(void* h);
void orc_get_contacts(void* h, long long* shape_pair, int* body_pair2, double* normal3, double* depth,
                      double* pt1, double* pt2, double* erad);
void orc_get_contact_forces(void* h, double* force3_on_b2, double* torque1_loc, double* torque2_loc);
void orc_get_body_forces(void* h, double* force3, double* torque3);
long long orc_num_history(void* h);
void orc_get_history(void* h, int* body, int* other, int* shape1, int* shape2, double* disp3, double* duration,
                     double* relvel_init);
void orc_add_history(void* h, int body, int other, int shape1, int shape2, const double disp[3], double duration,
                     double relvel_init);
/* seconds accumulated: [0]=aabb+broadphase [1]=narrowphase [2]=materials+forces+reduce [3]=integrate [4]=total */
void orc_get_timers(void* h, double t[5]);
void orc_reset_timers(void* h);
int orc_max_threads(void);

/* Narrowphase primitives, for the golden vectors (return 1 on contact).
 * src/chrono/collision/multicore/ChNarrowphasePRIMS.cpp:40-72, 269-313, 379-437;
 * ChCollisionUtils.h:546-563; ChCollisionUtilsPRIMS.cpp:41-106 */
int orc_sphere_sphere(const double pos1[3], double r1, const double pos2[3], double r2, double separation,
                      double norm[3], double* depth, double pt1[3], double pt2[3], double* erad);
int orc_box_sphere(const double pos1[3], const double rot1[4], const double hdims1[3], const double pos2[3],
                   double r2, double separation, double norm[3], double* depth, double pt1[3], double pt2[3],
                   double* erad);
int orc_triangle_sphere(const double A[3], const double B[3], const double C[3], const double pos2[3], double r2,
                        double separation, double norm[3], double* depth, double pt1[3], double pt2[3],
                        double* erad);
unsigned orc_snap_to_box(const double hdims[3], double loc[3]);
int orc_snap_to_triangle(const double A[3], const double B[3], const double C[3], const double P[3],
                         double res[3]);
/* Rotate / RotateT / AbsRotate (src/chrono/multicore_math/real4.cpp:158-187) */
void orc_rotate(const double v[3], const double q[4], double out[3], double outT[3], double outAbs[3]);
/* composite material, src/chrono/physics/ChContactMaterialSMC.cpp:107-130; out[13] =
 * E_eff,G_eff,mu,muRoll,muSpin,cr,adh,adhDMT,adhPerko,kn,kt,gn,gt (floats widened to double) */
void orc_composite(const OrcMaterial* m1, const OrcMaterial* m2, double out[13]);
/* One contact through the force law (function_CalcContactForces,
 * src/chrono_multicore/solver/ChIterativeSolverMulticoreSMC.cpp:56-546) for a two-body system with body ids
 * b1,b2 in {0,1} and shape ids s1=b1, s2=b2.  mass[2], pos[2][3], rot[2][4], vel[2][6] are indexed by body id.
 * hist_*: in/out history of the pair stored on body max(b1,b2) (MultiStep only); hist_present=0 -> no entry yet.
 * Returns 0, or 1 if no free history slot. */
int orc_contact_force(const OrcSettings* s, const double comp[13], int b1, int b2, const double* mass,
                      const double* pos, const double* rot, const double* vel, const double pt1[3],
                      const double pt2[3], const double normal[3], double depth, double erad, int* hist_present,
                      double hist_disp[3], double* hist_dur, double* hist_relvel, double force_b2[3],
                      double torque_b1[3], double torque_b2[3]);

#ifdef __cplusplus
}
#endif
#endif

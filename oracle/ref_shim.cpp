// =============================================================================
// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points into the REFERENCE's own
// code, compiled (unmodified, from /root/reference, by oracle/Makefile) into
// oracle/_ref/libchrono_ref.so.  Nothing here restates reference arithmetic: every function only
// marshals plain arrays into the reference's data structures and calls the reference symbol.
// Used by tests/test_oracle_vs_ref.py to pin the restatement in oracle/dem_oracle.cpp bit-for-bit.
// =============================================================================
#include "chrono/collision/multicore/ChBroadphase.h"
#include "chrono/collision/multicore/ChNarrowphase.h"
#include "chrono/collision/multicore/ChCollisionUtils.h"
#include "chrono/multicore_math/utility.h"
#include "smc_prelude.h"

#include <cstring>

namespace chrono {
// Defined in the reference's ChNarrowphasePRIMS.cpp (external linkage, no header declaration).
bool sphere_sphere(const real3& pos1, const real& radius1, const real3& pos2, const real& radius2,
                   const real& separation, real3& norm, real& depth, real3& pt1, real3& pt2, real& eff_radius);
bool box_sphere(const real3& pos1, const quaternion& rot1, const real3& hdims1, const real3& pos2,
                const real& radius2, const real& separation, real3& norm, real& depth, real3& pt1, real3& pt2,
                real& eff_radius);
bool triangle_sphere(const real3& A1, const real3& B1, const real3& C1, const real3& pos2, const real& radius2,
                     const real& separation, real3& norm, real& depth, real3& pt1, real3& pt2, real& eff_radius);

// The reference grants friend access to a class of this name (ChBroadphase.h:60, ChNarrowphase.h:196).  The real
// one lives in ChCollisionSystemMulticore.cpp, which needs ChSystem (-> Eigen) and is not compiled here; this
// stand-in only wires a ChCollisionData into the reference's broadphase / narrowphase objects and runs them.
class ChCollisionSystemMulticore {
  public:
    static void Run(std::shared_ptr<ChCollisionData> cd, const int bins[3]) {
        ChBroadphase bp;
        bp.cd_data = cd;
        bp.grid_type = ChBroadphase::GridType::FIXED_RESOLUTION;
        bp.grid_resolution = vec3(bins[0], bins[1], bins[2]);
        ChNarrowphase np;
        np.cd_data = cd;
        np.algorithm = ChNarrowphase::Algorithm::PRIMS;
        bp.Process();
        np.Process();
    }
};
}  // namespace chrono

using namespace chrono;

// Defined by the excerpt compiled from ChIterativeSolverMulticoreSMC.cpp:56-546 (see Makefile, smc_prelude.h)
void function_CalcContactForces(int index, vec2* body_pairs, vec2* shape_pairs,
                                ChSystemSMC::ContactForceModel contact_model,
                                ChSystemSMC::AdhesionForceModel adhesion_model,
                                ChSystemSMC::TangentialDisplacementModel displ_mode, bool use_mat_props, real char_vel, real min_slip_vel,
                                real min_roll_vel, real min_spin_vel, real dT, real* body_mass, real3* pos,
                                quaternion* rot, real* vel, real3* friction, real2* modulus, real3* adhesion,
                                real* cr, real4* smc_params, real3* pt1, real3* pt2, real3* normal, real* depth,
                                real* eff_radius, vec3* shear_neigh, char* shear_touch, real3* shear_disp,
                                real* contact_relvel_init, real* contact_duration, int* ct_bid, real3* ct_force,
                                real3* ct_torque);

static real3 r3(const double* p) { return real3(p[0], p[1], p[2]); }
static void w3(const real3& v, double* p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

extern "C" {

int ref_sphere_sphere(const double pos1[3], double r1, const double pos2[3], double r2, double separation,
                      double norm[3], double* depth, double pt1[3], double pt2[3], double* erad) {
    real3 n, p1, p2;
    real d = 0, e = 0;
    bool hit = sphere_sphere(r3(pos1), r1, r3(pos2), r2, separation, n, d, p1, p2, e);
    if (hit) { w3(n, norm); w3(p1, pt1); w3(p2, pt2); *depth = d; *erad = e; }
    return hit;
}
int ref_box_sphere(const double pos1[3], const double rot1[4], const double hdims1[3], const double pos2[3],
                   double r2, double separation, double norm[3], double* depth, double pt1[3], double pt2[3],
                   double* erad) {
    real3 n, p1, p2;
    real d = 0, e = 0;
    bool hit = box_sphere(r3(pos1), quaternion(rot1[0], rot1[1], rot1[2], rot1[3]), r3(hdims1), r3(pos2), r2,
                          separation, n, d, p1, p2, e);
    if (hit) { w3(n, norm); w3(p1, pt1); w3(p2, pt2); *depth = d; *erad = e; }
    return hit;
}
int ref_triangle_sphere(const double A[3], const double B[3], const double C[3], const double pos2[3], double r2,
                        double separation, double norm[3], double* depth, double pt1[3], double pt2[3],
                        double* erad) {
    real3 n, p1, p2;
    real d = 0, e = 0;
    bool hit = triangle_sphere(r3(A), r3(B), r3(C), r3(pos2), r2, separation, n, d, p1, p2, e);
    if (hit) { w3(n, norm); w3(p1, pt1); w3(p2, pt2); *depth = d; *erad = e; }
    return hit;
}
unsigned ref_snap_to_box(const double hdims[3], double loc[3]) {
    real3 l = r3(loc);
    unsigned code = mc_utils::snap_to_box(r3(hdims), l);
    w3(l, loc);
    return code;
}
int ref_snap_to_triangle(const double A[3], const double B[3], const double C[3], const double P[3], double res[3]) {
    real3 r;
    bool edge = mc_utils::snap_to_triangle(r3(A), r3(B), r3(C), r3(P), r);
    w3(r, res);
    return edge;
}
void ref_rotate(const double v[3], const double q[4], double out[3], double outT[3], double outAbs[3]) {
    quaternion Q(q[0], q[1], q[2], q[3]);
    w3(Rotate(r3(v), Q), out);
    w3(RotateT(r3(v), Q), outT);
    w3(AbsRotate(Q, r3(v)), outAbs);
}

// Reference broadphase + narrowphase on a caller-described scene.
//   shapes: type[ns] (ChCollisionShape::Type), body[ns], lpos[3ns], lrot[4ns], dims[3ns] (sphere r,0,0 / box hdims),
//           tri[9ns] (body frame; ignored unless TRIANGLE)
//   bodies: pos[3nb], rot[4nb], active[nb], collide[nb]
//   aabb_min/max[3ns]: shape AABBs (world frame, not yet offset), as GenerateAABB would produce them
// Results are kept in a static ChCollisionData and read back with the ref_cd_* getters.
static std::shared_ptr<ChCollisionData> g_cd;

void ref_collision_run(int ns, const int* type, const int* body, const double* lpos, const double* lrot,
                       const double* dims, const double* tri, int nb, const double* pos, const double* rot,
                       const char* active, const char* collide, const double* aabb_min, const double* aabb_max,
                       const int bins[3]) {
    g_cd = std::make_shared<ChCollisionData>(true);
    ChCollisionData& cd = *g_cd;
    shape_container& sd = cd.shape_data;
    for (int i = 0; i < ns; i++) {
        sd.fam_rigid.push_back(short2(1, 0x7FFF));
        sd.id_rigid.push_back(body[i]);
        sd.typ_rigid.push_back(type[i]);
        sd.local_rigid.push_back(0);
        sd.length_rigid.push_back(1);
        sd.ObA_rigid.push_back(r3(lpos + 3 * i));
        sd.ObR_rigid.push_back(quaternion(lrot[4 * i], lrot[4 * i + 1], lrot[4 * i + 2], lrot[4 * i + 3]));
        if (type[i] == ChCollisionShape::Type::SPHERE) {
            sd.start_rigid.push_back((int)sd.sphere_rigid.size());
            sd.sphere_rigid.push_back(dims[3 * i]);
        } else if (type[i] == ChCollisionShape::Type::BOX) {
            sd.start_rigid.push_back((int)sd.box_like_rigid.size());
            sd.box_like_rigid.push_back(r3(dims + 3 * i));
        } else {
            sd.start_rigid.push_back((int)sd.triangle_rigid.size());
            for (int k = 0; k < 3; k++)
                sd.triangle_rigid.push_back(r3(tri + 9 * i + 3 * k));
        }
        cd.aabb_min.push_back(r3(aabb_min + 3 * i));
        cd.aabb_max.push_back(r3(aabb_max + 3 * i));
    }
    cd.num_rigid_shapes = ns;
    cd.state_data.num_rigid_bodies = nb;
    for (int i = 0; i < nb; i++) {
        cd.state_data.pos_rigid->push_back(r3(pos + 3 * i));
        cd.state_data.rot_rigid->push_back(quaternion(rot[4 * i], rot[4 * i + 1], rot[4 * i + 2], rot[4 * i + 3]));
        cd.state_data.active_rigid->push_back(active[i]);
        cd.state_data.collide_rigid->push_back(collide[i]);
    }
    cd.collision_envelope = 0;
    cd.bins_per_axis = vec3(bins[0], bins[1], bins[2]);
    ChCollisionSystemMulticore::Run(g_cd, bins);
}
void ref_cd_sizes(long long s[4]) {
    s[0] = g_cd->num_active_bins;
    s[1] = g_cd->num_bin_aabb_intersections;
    s[2] = (long long)g_cd->pair_shapeIDs.size();
    s[3] = g_cd->num_rigid_contacts;
}
void ref_cd_grid(double origin[3], double bin_size[3], double inv_bin_size[3]) {
    w3(g_cd->global_origin, origin);
    w3(g_cd->bin_size, bin_size);
    w3(g_cd->inv_bin_size, inv_bin_size);
}
void ref_cd_bins(unsigned* bin_active, unsigned* bin_start_index, unsigned* bin_aabb_number) {
    std::memcpy(bin_active, g_cd->bin_active.data(), sizeof(unsigned) * g_cd->num_active_bins);
    std::memcpy(bin_start_index, g_cd->bin_start_index.data(), sizeof(unsigned) * (g_cd->num_active_bins + 1));
    std::memcpy(bin_aabb_number, g_cd->bin_aabb_number.data(), sizeof(unsigned) * g_cd->num_bin_aabb_intersections);
}
void ref_cd_pairs(long long* p) {
    std::memcpy(p, g_cd->pair_shapeIDs.data(), sizeof(long long) * g_cd->pair_shapeIDs.size());
}
void ref_cd_contacts(long long* shape_pair, int* body_pair2, double* normal3, double* depth, double* pt1,
                     double* pt2, double* erad) {
    for (unsigned i = 0; i < g_cd->num_rigid_contacts; i++) {
        shape_pair[i] = g_cd->contact_shapeIDs[i];
        body_pair2[2 * i] = g_cd->bids_rigid_rigid[i].x;
        body_pair2[2 * i + 1] = g_cd->bids_rigid_rigid[i].y;
        w3(g_cd->norm_rigid_rigid[i], normal3 + 3 * i);
        depth[i] = g_cd->dpth_rigid_rigid[i];
        w3(g_cd->cpta_rigid_rigid[i], pt1 + 3 * i);
        w3(g_cd->cptb_rigid_rigid[i], pt2 + 3 * i);
        erad[i] = g_cd->erad_rigid_rigid[i];
    }
}

// Same contract as orc_contact_force (oracle/dem_oracle.h), evaluated by the reference's function_CalcContactForces.
int ref_contact_force(const int* model4 /*force,adhesion,tang,use_mat*/, const double* par5 /*char_vel,slip,roll,spin,dT*/,
                      const double comp[13], int b1, int b2, const double* mass, const double* pos, const double* rot,
                      const double* vel, const double pt1[3], const double pt2[3], const double normal[3], double depth,
                      double erad, int* hist_present, double hist_disp[3], double* hist_dur, double* hist_relvel,
                      double force_b2[3], double torque_b1[3], double torque_b2[3]) {
    vec2 bp(b1, b2), sp(b1, b2);
    real m[2] = {mass[0], mass[1]};
    real3 p[2] = {r3(pos), r3(pos + 3)};
    quaternion q[2] = {quaternion(rot[0], rot[1], rot[2], rot[3]), quaternion(rot[4], rot[5], rot[6], rot[7])};
    real v[12];
    for (int i = 0; i < 12; i++) v[i] = vel[i];
    real3 fric(comp[2], comp[3], comp[4]);
    real2 mod(comp[0], comp[1]);
    real3 adh(comp[6], comp[7], comp[8]);
    real cr = comp[5];
    real4 smc(comp[9], comp[10], comp[11], comp[12]);
    real3 P1 = r3(pt1), P2 = r3(pt2), N = r3(normal);
    real dep = depth, er = erad;
    vec3 neigh[40];
    char touch[40];
    real3 disp[40];
    real relv[40], dur[40];
    for (int i = 0; i < 40; i++) {
        neigh[i] = vec3(-1, -1, -1);
        touch[i] = 0;
        disp[i] = real3(0);
        relv[i] = 0;
        dur[i] = 0;
    }
    int owner = b1 > b2 ? b1 : b2, other = b1 > b2 ? b2 : b1;
    if (*hist_present) {
        neigh[20 * owner] = vec3(other, owner, other);
        disp[20 * owner] = r3(hist_disp);
        relv[20 * owner] = *hist_relvel;
        dur[20 * owner] = *hist_dur;
    }
    int bid[2];
    real3 F[2], T[2];
    function_CalcContactForces(0, &bp, &sp, (ChSystemSMC::ContactForceModel)model4[0],
                               (ChSystemSMC::AdhesionForceModel)model4[1],
                               (ChSystemSMC::TangentialDisplacementModel)model4[2], model4[3] != 0, par5[0], par5[1], par5[2],
                               par5[3], par5[4], m, p, q, v, &fric, &mod, &adh, &cr, &smc, &P1, &P2, &N, &dep, &er,
                               neigh, touch, disp, relv, dur, bid, F, T);
    if (model4[2] == 2 && depth < 0) {
        *hist_present = 1;
        w3(disp[20 * owner], hist_disp);
        *hist_dur = dur[20 * owner];
        *hist_relvel = relv[20 * owner];
    }
    w3(F[1], force_b2);
    w3(T[0], torque_b1);
    w3(T[1], torque_b2);
    return 0;
}

}  // extern "C"

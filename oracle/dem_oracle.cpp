// =============================================================================
// oracle/dem_oracle.cpp -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
//
// fp64 restatement of one Chrono::Multicore SMC step, ChSystemMulticore::AdvanceDynamics
// (src/chrono_multicore/physics/ChSystemMulticore.cpp:73-200), restricted to sphere / box / triangle
// collision shapes, no bilateral constraints:
//   Update()                  ChSystemMulticore.cpp:302-377, ChBody.cpp:247-256,552-555 (hf = h*(m g), gyro)
//   GenerateAABB              src/chrono/collision/multicore/ChCollisionSystemMulticore.cpp:376-406,458-545
//   ChBroadphase::Process     ChBroadphase.cpp:85-102,143-345; ChCollisionUtils.h:44-110;
//                             ChCollisionUtilsBroadphase.cpp:51-217
//   ChNarrowphase (PRIMS)     ChNarrowphase.cpp:94-155,157-196,325-386; ChNarrowphasePRIMS.cpp:40-72,269-313,
//                             379-437,1406-1531; ChCollisionUtils.h:467-474,546-563;
//                             ChCollisionUtilsPRIMS.cpp:41-106
//   composite materials       ChContactContainerMulticoreSMC.cpp:119-166, ChContactMaterialSMC.cpp:107-130,
//                             ChContactMaterial.h:157-169  (evaluated in float, as the reference does)
//   contact forces            ChIterativeSolverMulticoreSMC.cpp:56-546 (function_CalcContactForces)
//   per-body reduction        ChIterativeSolverMulticoreSMC.cpp:644-725, 604-619
//   velocity update           ChIterativeSolverMulticore.cpp:60-159 (M_invk = v + M^-1 hf), :839-853
//   position update           ChSystemMulticore.cpp:131-152, ChBody.cpp:263-310
//
// Deliberate, documented deviations from the reference *implementation* (not its arithmetic):
//   * thrust::sort_by_key calls are replaced by std::stable_sort (the reference order of equal keys is
//     implementation-defined; every consumer of this oracle compares order-independent sets / sums).
//   * history slots are assigned in a serial pre-pass, in contact order, instead of inside the
//     OpenMP contact loop (the reference races on free slots, ChIterativeSolverMulticoreSMC.cpp:211-226).
//   * per-body force sums are accumulated in contact order (reference: order after an unstable sort).
// Everything else keeps the reference's operation order; build with -ffp-contract=off.
// =============================================================================
#include "dem_oracle.h"
#include "omath.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstring>
#include <parallel/algorithm>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

constexpr double kPI = 3.141592653589793238462643383279;  // ChConstants.h:18
constexpr double k2_3 = 2.0 / 3.0;                        // ChConstants.h:32
constexpr int kMaxShear = 20;                             // ChDataManager.h:185
constexpr double kEdgeRadius = 0.1;                       // ChNarrowphasePRIMS.cpp:31

enum ShapeType { SPHERE = 0, BOX = 2, TRIANGLE = 13 };  // ChCollisionShape.h:31-53

struct I3 {
    int x, y, z;
};

// ---------------------------------------------------------------------------------------------
// Narrowphase primitives
// ---------------------------------------------------------------------------------------------

// ChNarrowphasePRIMS.cpp:40-72
bool sphere_sphere(const V3& pos1, double radius1, const V3& pos2, double radius2, double separation, V3& norm,
                   double& depth, V3& pt1, V3& pt2, double& eff_radius) {
    V3 delta = pos2 - pos1;
    double dist2 = Dot(delta, delta);
    double radSum = radius1 + radius2;
    double radSum_s = radSum + separation;
    if (dist2 >= radSum_s * radSum_s || dist2 < 1e-12)
        return false;
    double dist = std::sqrt(dist2);
    norm = delta / dist;
    pt1 = pos1 + norm * radius1;
    pt2 = pos2 - norm * radius2;
    depth = dist - radSum;
    eff_radius = radius1 * radius2 / radSum;
    return true;
}

// ChCollisionUtils.h:546-563
unsigned snap_to_box(const V3& hdims, V3& loc) {
    unsigned code = 0;
    if (std::abs(loc.x) > hdims.x) {
        code |= 1;
        loc.x = (loc.x > 0) ? hdims.x : -hdims.x;
    }
    if (std::abs(loc.y) > hdims.y) {
        code |= 2;
        loc.y = (loc.y > 0) ? hdims.y : -hdims.y;
    }
    if (std::abs(loc.z) > hdims.z) {
        code |= 4;
        loc.z = (loc.z > 0) ? hdims.z : -hdims.z;
    }
    return code;
}

// ChNarrowphasePRIMS.cpp:269-313
bool box_sphere(const V3& pos1, const Q4& rot1, const V3& hdims1, const V3& pos2, double radius2, double separation,
                V3& norm, double& depth, V3& pt1, V3& pt2, double& eff_radius) {
    V3 spherePos = TransformParentToLocal(pos1, rot1, pos2);
    V3 boxPos = spherePos;
    unsigned code = snap_to_box(hdims1, boxPos);
    V3 delta = spherePos - boxPos;
    double dist2 = Dot(delta, delta);
    double radius2_s = radius2 + separation;
    if (dist2 >= radius2_s * radius2_s || dist2 <= 1e-12f)  // note: float literal in the reference
        return false;
    double dist = std::sqrt(dist2);
    depth = dist - radius2;
    norm = Rotate(delta / dist, rot1);
    pt1 = TransformLocalToParent(pos1, rot1, boxPos);
    pt2 = pos2 - norm * radius2;
    if ((code != 1) && (code != 2) && (code != 4))
        eff_radius = radius2 * kEdgeRadius / (radius2 + kEdgeRadius);
    else
        eff_radius = radius2;
    return true;
}

// ChCollisionUtils.h:467-474
V3 triangle_normal(const V3& A, const V3& B, const V3& C) {
    V3 v1 = B - A;
    V3 v2 = C - A;
    V3 n = Cross(v1, v2);
    double len = Length(n);
    return n / len;
}

// ChCollisionUtilsPRIMS.cpp:41-106 (Ericson, Real-time collision detection, p.141)
bool snap_to_triangle(const V3& A, const V3& B, const V3& C, const V3& P, V3& res) {
    V3 AB = B - A;
    V3 AC = C - A;
    V3 AP = P - A;
    double d1 = Dot(AB, AP);
    double d2 = Dot(AC, AP);
    if (d1 <= 0 && d2 <= 0) {
        res = A;
        return true;
    }
    V3 BP = P - B;
    double d3 = Dot(AB, BP);
    double d4 = Dot(AC, BP);
    if (d3 >= 0 && d4 <= d3) {
        res = B;
        return true;
    }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double v = d1 / (d1 - d3);
        res = A + v * AB;
        return true;
    }
    V3 CP = P - C;
    double d5 = Dot(AB, CP);
    double d6 = Dot(AC, CP);
    if (d6 >= 0 && d5 <= d6) {
        res = C;
        return true;
    }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double w = d2 / (d2 - d6);
        res = A + w * AC;
        return true;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        res = B + w * (C - B);
        return true;
    }
    double denom = 1 / (va + vb + vc);
    double v = vb * denom;
    double w = vc * denom;
    res = A + v * AB + w * AC;
    return false;
}

// ChNarrowphasePRIMS.cpp:379-437
bool triangle_sphere(const V3& A1, const V3& B1, const V3& C1, const V3& pos2, double radius2, double separation,
                     V3& norm, double& depth, V3& pt1, V3& pt2, double& eff_radius) {
    double radius2_s = radius2 + separation;
    V3 nrm1 = triangle_normal(A1, B1, C1);
    double h = Dot(pos2 - A1, nrm1);
    if (h >= radius2_s || h <= 0)
        return false;
    V3 faceLoc;
    if (snap_to_triangle(A1, B1, C1, pos2, faceLoc)) {
        V3 delta = pos2 - faceLoc;
        double dist2 = Dot(delta, delta);
        if (dist2 >= radius2_s * radius2_s || dist2 <= 1e-12f)
            return false;
        double dist = std::sqrt(dist2);
        norm = delta / dist;
        depth = dist - radius2;
        eff_radius = radius2 * kEdgeRadius / (radius2 + kEdgeRadius);
    } else {
        norm = nrm1;
        depth = h - radius2;
        eff_radius = radius2;
    }
    pt1 = faceLoc;
    pt2 = pos2 - norm * radius2;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Composite material (float arithmetic, as in the reference)
// ---------------------------------------------------------------------------------------------
struct Composite {
    float E_eff, G_eff, mu_eff, muRoll_eff, muSpin_eff, cr_eff, adhesion_eff, adhesionMultDMT_eff, adhesionSPerko_eff;
    float kn, kt, gn, gt;
};

// ChContactMaterialSMC.cpp:107-130 with the default strategy ChContactMaterial.h:161-169
Composite make_composite(const OrcMaterial& m1, const OrcMaterial& m2) {
    Composite c;
    float inv_E = (1 - m1.poisson * m1.poisson) / m1.young + (1 - m2.poisson * m2.poisson) / m2.young;
    float inv_G = 2 * (2 - m1.poisson) * (1 + m1.poisson) / m1.young + 2 * (2 - m2.poisson) * (1 + m2.poisson) / m2.young;
    c.E_eff = 1 / inv_E;
    c.G_eff = 1 / inv_G;
    c.mu_eff = std::min<float>(m1.mu_s, m2.mu_s);
    c.muRoll_eff = std::min<float>(m1.mu_roll, m2.mu_roll);
    c.muSpin_eff = std::min<float>(m1.mu_spin, m2.mu_spin);
    c.cr_eff = std::min<float>(m1.cr, m2.cr);
    c.adhesion_eff = std::min<float>(m1.adhesion, m2.adhesion);
    c.adhesionMultDMT_eff = std::min<float>(m1.adhesion_dmt, m2.adhesion_dmt);
    c.adhesionSPerko_eff = std::min<float>(m1.adhesion_perko, m2.adhesion_perko);
    c.kn = (m1.kn + m2.kn) / 2;
    c.kt = (m1.kt + m2.kt) / 2;
    c.gn = (m1.gn + m2.gn) / 2;
    c.gt = (m1.gt + m2.gt) / 2;
    return c;
}

// ---------------------------------------------------------------------------------------------
// Contact force law: ChIterativeSolverMulticoreSMC.cpp:56-546.  Array arguments as in the reference.
// The history slot for MultiStep is located by the caller (contact_id), see header comment.
// ---------------------------------------------------------------------------------------------
struct ForceIn {
    int force_model, adhesion_model, displ_mode;
    bool use_mat_props;
    double char_vel, min_slip_vel, min_roll_vel, min_spin_vel, dT;
};

struct HistSlot {  // one slot of shear_neigh/shear_disp/contact_relvel_init/contact_duration
    int nb, s1, s2;  // shear_neigh: other body, larger shape id, smaller shape id; nb=-1 -> free
    V3 disp;
    double relvel_init, duration;
    char touch;
};

void calc_contact_force(const ForceIn& P, int b1, int b2, const double* body_mass, const V3* pos, const Q4* rot,
                        const double* vel, const Composite& cm, const V3& pt1, const V3& pt2, const V3& normal,
                        double depth, double eff_radius, HistSlot* slot /* may be null */, bool newcontact,
                        V3& out_force /* on b2; -force on b1 */, V3& out_torque1, V3& out_torque2) {
    if (depth >= 0) {  // :96-104
        out_force = V3(0);
        out_torque1 = V3(0);
        out_torque2 = V3(0);
        return;
    }
    V3 pt1_loc = TransformParentToLocal(pos[b1], rot[b1], pt1);
    V3 pt2_loc = TransformParentToLocal(pos[b2], rot[b2], pt2);
    V3 v_body1(vel[b1 * 6 + 0], vel[b1 * 6 + 1], vel[b1 * 6 + 2]);
    V3 v_body2(vel[b2 * 6 + 0], vel[b2 * 6 + 1], vel[b2 * 6 + 2]);
    V3 o_body1(vel[b1 * 6 + 3], vel[b1 * 6 + 4], vel[b1 * 6 + 5]);
    V3 o_body2(vel[b2 * 6 + 3], vel[b2 * 6 + 4], vel[b2 * 6 + 5]);
    V3 vel1 = v_body1 + Rotate(Cross(o_body1, pt1_loc), rot[b1]);
    V3 vel2 = v_body2 + Rotate(Cross(o_body2, pt2_loc), rot[b2]);
    V3 relvel = vel2 - vel1;
    double relvel_n_mag = Dot(relvel, normal);
    V3 relvel_n = relvel_n_mag * normal;
    V3 relvel_t = relvel - relvel_n;
    double relvel_t_mag = Length(relvel_t);

    double m_eff = body_mass[b1] * body_mass[b2] / (body_mass[b1] + body_mass[b2]);
    double mu_eff = cm.mu_eff, muRoll_eff = cm.muRoll_eff, muSpin_eff = cm.muSpin_eff;
    double E_eff = cm.E_eff, G_eff = cm.G_eff;
    double adhesion_eff = cm.adhesion_eff, adhesionMultDMT_eff = cm.adhesionMultDMT_eff,
           adhesionSPerko_eff = cm.adhesionSPerko_eff;
    double cr_eff = cm.cr_eff;
    double user_kn = cm.kn, user_kt = cm.kt, user_gn = cm.gn, user_gt = cm.gt;

    double kn = 0, kt = 0, gn = 0, gt = 0, kn_simple = 0, gn_simple = 0;
    double t_contact = 0;
    double relvel_init = std::abs(relvel_n_mag);
    double delta_n = -depth;
    V3 delta_t(0);
    double char_vel = P.char_vel;
    bool hist_on_b1 = (std::max(b1, b2) == b1);

    if (P.displ_mode == ORC_TANG_ONESTEP) {
        delta_t = relvel_t * P.dT;
    } else if (P.displ_mode == ORC_TANG_MULTISTEP) {
        delta_t = relvel_t * P.dT;
        // :201-227 (slot lookup done by the caller; same state transitions)
        if (!newcontact) {
            slot->duration += P.dT;
        } else {
            slot->disp = V3(0);
            slot->relvel_init = relvel_init;
            slot->duration = 0;
        }
        slot->touch = 1;
        // :233-243
        if (hist_on_b1) {
            slot->disp += delta_t;
            slot->disp -= Dot(slot->disp, normal) * normal;
            delta_t = slot->disp;
        } else {
            slot->disp -= delta_t;
            slot->disp -= Dot(slot->disp, normal) * normal;
            delta_t = -slot->disp;
        }
        relvel_init = (slot->relvel_init < char_vel) ? char_vel : slot->relvel_init;
        t_contact = slot->duration;
    }

    double eps = std::numeric_limits<double>::epsilon();

    switch (P.force_model) {
        case ORC_HOOKE:
            if (P.use_mat_props) {
                double tmp_k = (16.0 / 15) * std::sqrt(eff_radius) * E_eff;
                char_vel = (P.displ_mode == ORC_TANG_MULTISTEP) ? relvel_init : char_vel;
                double v2 = char_vel * char_vel;
                double loge = (cr_eff < eps) ? std::log(eps) : std::log(cr_eff);
                loge = (cr_eff > 1 - eps) ? std::log(1 - eps) : loge;
                double tmp_g = 1 + std::pow(kPI / loge, 2);
                kn = tmp_k * std::pow(m_eff * v2 / tmp_k, 0.2);
                kt = kn;
                gn = std::sqrt(4 * m_eff * kn / tmp_g);
                gt = gn;
            } else {
                kn = user_kn;
                kt = user_kt;
                gn = m_eff * user_gn;
                gt = m_eff * user_gt;
            }
            kn_simple = kn;
            gn_simple = gn;
            break;
        case ORC_HERTZ:
            if (P.use_mat_props) {
                double sqrt_Rd = std::sqrt(eff_radius * delta_n);
                double Sn = 2 * E_eff * sqrt_Rd;
                double St = 8 * G_eff * sqrt_Rd;
                double loge = (cr_eff < eps) ? std::log(eps) : std::log(cr_eff);
                double beta = loge / std::sqrt(loge * loge + kPI * kPI);
                kn = k2_3 * Sn;
                kt = St;
                gn = -2 * std::sqrt(5.0 / 6) * beta * std::sqrt(Sn * m_eff);
                gt = -2 * std::sqrt(5.0 / 6) * beta * std::sqrt(St * m_eff);
            } else {
                double tmp = eff_radius * std::sqrt(delta_n);
                kn = tmp * user_kn;
                kt = tmp * user_kt;
                gn = tmp * m_eff * user_gn;
                gt = tmp * m_eff * user_gt;
            }
            kn_simple = kn / std::sqrt(delta_n);
            gn_simple = gn / std::pow(delta_n, 0.25);
            break;
        case ORC_FLORES:
            if (P.use_mat_props) {
                double sqrt_Rd = std::sqrt(eff_radius * delta_n);
                double Sn = 2 * E_eff * sqrt_Rd;
                double St = 8 * G_eff * sqrt_Rd;
                cr_eff = (cr_eff < 0.01) ? 0.01 : cr_eff;
                cr_eff = (cr_eff > 1.0 - eps) ? 1.0 - eps : cr_eff;
                double loge = std::log(cr_eff);
                double beta = loge / std::sqrt(loge * loge + kPI * kPI);
                char_vel = (P.displ_mode == ORC_TANG_MULTISTEP) ? relvel_init : char_vel;
                kn = k2_3 * Sn;
                kt = k2_3 * St;
                gn = 8.0 * (1.0 - cr_eff) * kn * delta_n / (5.0 * cr_eff * char_vel);
                gt = -2 * std::sqrt(5.0 / 6) * beta * std::sqrt(St * m_eff);
            } else {
                double tmp = eff_radius * std::sqrt(delta_n);
                kn = tmp * user_kn;
                kt = tmp * user_kt;
                gn = tmp * m_eff * user_gn * delta_n;
                gt = tmp * m_eff * user_gt;
            }
            kn_simple = kn / std::sqrt(delta_n);
            gn_simple = gn / std::pow(delta_n, 1.5);
            break;
        case ORC_PLAINCOULOMB:
            if (P.use_mat_props) {
                double sqrt_Rd = std::sqrt(delta_n);
                double Sn = 2 * E_eff * sqrt_Rd;
                double loge = (cr_eff < eps) ? std::log(eps) : std::log(cr_eff);
                double beta = loge / std::sqrt(loge * loge + kPI * kPI);
                kn = k2_3 * Sn;
                gn = -2 * std::sqrt(5.0 / 6) * beta * std::sqrt(Sn * m_eff);
            } else {
                double tmp = std::sqrt(delta_n);
                kn = tmp * user_kn;
                gn = tmp * user_gn;
            }
            kn_simple = kn / std::sqrt(delta_n);
            gn_simple = gn / std::pow(delta_n, 0.25);
            kt = 0;
            gt = 0;
            break;
    }

    double forceN_mag = kn * delta_n - gn * relvel_n_mag;
    V3 force;
    if (P.force_model == ORC_PLAINCOULOMB) {  // :356-361
        double forceT_mag = mu_eff * std::tanh(5.0 * relvel_t_mag) * forceN_mag;
        force = forceN_mag * normal;
        if (relvel_t_mag >= P.min_slip_vel)
            force -= (forceT_mag / relvel_t_mag) * relvel_t;
    } else {  // :437-471
        V3 forceT_stiff = kt * delta_t;
        V3 forceT_damp = gt * relvel_t;
        V3 forceT = forceT_stiff + forceT_damp;
        double forceT_mag = Length(forceT);
        double delta_t_mag = Length(delta_t);
        double forceT_slide = mu_eff * std::abs(forceN_mag);
        if (forceT_mag > forceT_slide) {
            if (delta_t_mag > eps) {
                double ratio = forceT_slide / forceT_mag;
                forceT *= ratio;
                if (P.displ_mode == ORC_TANG_MULTISTEP) {
                    delta_t = (forceT - forceT_damp) / kt;
                    if (hist_on_b1)
                        slot->disp = delta_t;
                    else
                        slot->disp = -delta_t;
                }
            } else {
                forceT = V3(0);
            }
        }
        force = forceN_mag * normal - forceT;
    }

    // :478-479 / :364-365
    V3 torque1_loc = Cross(pt1_loc, RotateT(force, rot[b1]));
    V3 torque2_loc = Cross(pt2_loc, RotateT(force, rot[b2]));

    // :484-491
    double d_coeff = gn_simple / (2.0 * m_eff * std::sqrt(kn_simple / m_eff));
    if (d_coeff < 1.0) {
        double t_collision = kPI * std::sqrt(m_eff / (kn_simple * (1 - d_coeff * d_coeff)));
        if (t_contact <= t_collision) {
            muRoll_eff = 0.0;
            muSpin_eff = 0.0;
        }
    }
    // :494-524
    V3 v_rot = Rotate(Cross(o_body2, pt2_loc), rot[b2]) - Rotate(Cross(o_body1, pt1_loc), rot[b1]);
    V3 rel_o = Rotate(o_body2, rot[b2]) - Rotate(o_body1, rot[b1]);
    V3 m_roll1(0), m_roll2(0);
    if (Length(v_rot) > P.min_roll_vel && muRoll_eff > eps) {
        m_roll1 = muRoll_eff * Cross(forceN_mag * pt1_loc, RotateT(v_rot, rot[b1])) / Length(v_rot);
        m_roll2 = muRoll_eff * Cross(forceN_mag * pt2_loc, RotateT(v_rot, rot[b2])) / Length(v_rot);
    }
    V3 m_spin1(0), m_spin2(0);
    if (Length(rel_o) > P.min_spin_vel && muSpin_eff > eps) {
        double r1 = Length(pt1_loc);
        double r2 = Length(pt2_loc);
        double xc = (r1 * r1 - r2 * r2) / (2 * (r1 + r2 - delta_n)) + 0.5 * (r1 + r2 - delta_n);
        double rc = r1 * r1 - xc * xc;
        rc = (rc < eps) ? eps : std::sqrt(rc);
        m_spin1 = muSpin_eff * rc * RotateT(Dot(rel_o, forceN_mag * normal) * normal, rot[b1]) / Length(rel_o);
        m_spin2 = muSpin_eff * rc * RotateT(Dot(rel_o, forceN_mag * normal) * normal, rot[b2]) / Length(rel_o);
    }
    // :527-537
    switch (P.adhesion_model) {
        case ORC_ADH_CONSTANT:
            force -= adhesion_eff * normal;
            break;
        case ORC_ADH_DMT:
            force -= adhesionMultDMT_eff * std::sqrt(eff_radius) * normal;
            break;
        case ORC_ADH_PERKO:
            force -= adhesionSPerko_eff * eff_radius * normal;
            break;
    }
    // :540-545
    out_force = force;
    out_torque1 = -torque1_loc + m_roll1 + m_spin1;
    out_torque2 = torque2_loc - m_roll2 - m_spin2;
}

// ---------------------------------------------------------------------------------------------
// System
// ---------------------------------------------------------------------------------------------
struct Shape {
    int type, body, material;
    V3 A;        // local position   (ObA_rigid)
    Q4 R;        // local rotation   (ObR_rigid)
    V3 dims;     // sphere: (r,0,0); box: half dims
    V3 tri[3];   // triangle vertices, body frame
};

struct System {
    OrcSettings st;
    std::vector<OrcMaterial> mats;
    // bodies
    std::vector<double> mass;
    std::vector<V3> inv_inertia;  // diagonal of J = I^-1 (body frame)
    std::vector<V3> inertia;
    std::vector<V3> pos;
    std::vector<Q4> rot;
    std::vector<double> v;        // 6 per body: linear (global), angular (local)
    std::vector<char> active, collide;
    std::vector<V3> gyro;
    // shapes
    std::vector<Shape> shapes;
    // history
    std::vector<HistSlot> hist;  // kMaxShear per body
    // broadphase results
    std::vector<V3> aabb_min, aabb_max;
    V3 origin, bin_size, inv_bin_size, min_bp, max_bp;
    std::vector<I3> gmin, gmax;
    std::vector<unsigned> bin_intersections, bin_number, bin_aabb_number, bin_active, bin_start_index, bin_num_contact;
    unsigned num_active_bins = 0;
    std::vector<long long> pair_shapeIDs;
    // narrowphase results
    std::vector<V3> shapeA_global;
    std::vector<Q4> shapeR_global;
    std::vector<V3> tri_global;  // 3 per shape (unused for non-triangles)
    // narrowphase / force scratch (kept between steps)
    std::vector<char> np_act;
    std::vector<V3> np_norm, np_p1, np_p2;
    std::vector<double> np_dep, np_er;
    std::vector<long long> np_offs;
    std::vector<int> pc_slot_of;
    std::vector<char> pc_is_new;
    std::vector<long long> c_shape;
    std::vector<int> c_b1, c_b2;
    std::vector<V3> c_norm, c_pt1, c_pt2;
    std::vector<double> c_depth, c_erad;
    // forces
    std::vector<V3> c_force, c_tq1, c_tq2;
    std::vector<V3> body_force, body_torque;
    std::vector<double> hf;
    double timers[5] = {0, 0, 0, 0, 0};
};

inline double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ChCollisionUtils.h:44-65
inline I3 HashMin(const V3& A, const V3& inv) {
    return I3{(int)std::floor(A.x * inv.x), (int)std::floor(A.y * inv.y), (int)std::floor(A.z * inv.z)};
}
inline I3 HashMax(const V3& A, const V3& inv) {
    return I3{(int)std::ceil(A.x * inv.x) - 1, (int)std::ceil(A.y * inv.y) - 1, (int)std::ceil(A.z * inv.z) - 1};
}
inline unsigned Hash_Index(const I3& A, const int bins[3]) {
    return ((A.z * bins[1]) * bins[0]) + (A.y * bins[0]) + A.x;
}
// ChCollisionUtils.h:83-87
inline bool overlap(const V3& Amin, const V3& Amax, const V3& Bmin, const V3& Bmax) {
    return (Amin.x <= Bmax.x && Bmin.x <= Amax.x) && (Amin.y <= Bmax.y && Bmin.y <= Amax.y) &&
           (Amin.z <= Bmax.z && Bmin.z <= Amax.z);
}
// ChCollisionUtils.h:89-103
inline bool current_bin(const V3& Amin, const V3& Bmin, const V3& inv, const int bins[3], unsigned bin) {
    V3 min_p = Max(Amin, Bmin);
    return Hash_Index(HashMin(min_p, inv), bins) == bin;
}

// GenerateAABB: ChCollisionSystemMulticore.cpp:458-545 (collision envelope = 0 for SMC)
void generate_aabb(System& S) {
    const int ns = (int)S.shapes.size();
    S.aabb_min.resize(ns);
    S.aabb_max.resize(ns);
#pragma omp parallel for
    for (int i = 0; i < ns; i++) {
        const Shape& sh = S.shapes[i];
        const V3& position = S.pos[sh.body];
        const Q4& brot = S.rot[sh.body];
        V3 mn, mx;
        if (sh.type == SPHERE) {  // :376-385
            V3 p = Rotate(sh.A, brot) + position;
            mn = p - sh.dims.x;
            mx = p + sh.dims.x;
        } else if (sh.type == BOX) {  // :395-406
            Q4 rotation = Mult(brot, sh.R);
            V3 temp = AbsRotate(rotation, sh.dims);
            V3 p = Rotate(sh.A, brot) + position;
            mn = p - temp;
            mx = p + temp;
        } else {  // TRIANGLE :387-394, :524-535
            V3 A = Rotate(sh.tri[0], brot) + position;
            V3 B = Rotate(sh.tri[1], brot) + position;
            V3 C = Rotate(sh.tri[2], brot) + position;
            mn = V3(std::min(A.x, std::min(B.x, C.x)), std::min(A.y, std::min(B.y, C.y)), std::min(A.z, std::min(B.z, C.z)));
            mx = V3(std::max(A.x, std::max(B.x, C.x)), std::max(A.y, std::max(B.y, C.y)), std::max(A.z, std::max(B.z, C.z)));
        }
        S.aabb_min[i] = mn;
        S.aabb_max[i] = mx;
    }
}

// ChBroadphase::Process: ChBroadphase.cpp:213-226
void broadphase(System& S) {
    const int ns = (int)S.shapes.size();
    // RigidBoundingBox :85-102 (shapes of non-colliding bodies are excluded)
    V3 mn(std::numeric_limits<double>::max()), mx(-std::numeric_limits<double>::max());
    for (int i = 0; i < ns; i++) {
        if (!S.collide[S.shapes[i].body])
            continue;
        mn = Min(mn, S.aabb_min[i]);
        mx = Max(mx, S.aabb_max[i]);
    }
    // DetermineBoundingBox :143-166
    double fraction = 1e-3;
    V3 size = mx - mn;
    mn = mn - fraction * size;
    mx = mx + fraction * size;
    S.min_bp = mn;
    S.max_bp = mx;
    S.origin = mn;
    // OffsetAABB :168-176
#pragma omp parallel for
    for (int i = 0; i < ns; i++) {
        S.aabb_min[i] = S.aabb_min[i] - S.origin;
        S.aabb_max[i] = S.aabb_max[i] - S.origin;
    }
    // ComputeTopLevelResolution :179-208 (FIXED_RESOLUTION)
    const int* bins = S.st.bins_per_axis;
    V3 diag = Abs(S.max_bp - S.origin);
    S.bin_size = diag / V3(bins[0], bins[1], bins[2]);
    S.inv_bin_size = 1.0 / S.bin_size;
    const V3 inv = S.inv_bin_size;

    // OneLevelBroadphase :228-345
    S.gmin.resize(ns);
    S.gmax.resize(ns);
    S.bin_intersections.assign(ns + 1, 0);
#pragma omp parallel for
    for (int i = 0; i < ns; i++) {  // f_Count_AABB_BIN_Intersection
        I3 a = HashMin(S.aabb_min[i], inv);
        I3 b = HashMax(S.aabb_max[i], inv);
        S.gmin[i] = a;
        S.gmax[i] = b;
        S.bin_intersections[i] = (b.x - a.x + 1) * (b.y - a.y + 1) * (b.z - a.z + 1);
    }
    {  // Thrust_Exclusive_Scan
        unsigned run = 0;
        for (int i = 0; i <= ns; i++) {
            unsigned c = S.bin_intersections[i];
            S.bin_intersections[i] = run;
            run += c;
        }
    }
    const unsigned nint = S.bin_intersections[ns];
    S.bin_number.resize(nint);
    S.bin_aabb_number.resize(nint);
#pragma omp parallel for
    for (int index = 0; index < ns; index++) {  // f_Store_AABB_BIN_Intersection
        unsigned count = 0;
        I3 a = S.gmin[index], b = S.gmax[index];
        unsigned mInd = S.bin_intersections[index];
        for (int i = a.x; i <= b.x; i++)
            for (int j = a.y; j <= b.y; j++)
                for (int k = a.z; k <= b.z; k++) {
                    S.bin_number[mInd + count] = Hash_Index(I3{i, j, k}, bins);
                    S.bin_aabb_number[mInd + count] = index;
                    count++;
                }
    }
    {  // Thrust_Sort_By_Key(bin_number, bin_aabb_number) -> stable sort (see header note)
        std::vector<unsigned> perm(nint);
#pragma omp parallel for
        for (long long i = 0; i < (long long)nint; i++)
            perm[i] = (unsigned)i;
        // parallel (OpenMP) merge sort of libstdc++; stable, so the result equals std::stable_sort's
        __gnu_parallel::stable_sort(perm.begin(), perm.end(),
                                    [&](unsigned a, unsigned b) { return S.bin_number[a] < S.bin_number[b]; });
        std::vector<unsigned> k2(nint), v2(nint);
#pragma omp parallel for
        for (long long i = 0; i < (long long)nint; i++) {
            k2[i] = S.bin_number[perm[i]];
            v2[i] = S.bin_aabb_number[perm[i]];
        }
        S.bin_number.swap(k2);
        S.bin_aabb_number.swap(v2);
    }
    // Run_Length_Encode -> bin_active (unique bins), bin_start_index (counts -> exclusive scan)
    S.bin_active.clear();
    S.bin_start_index.clear();
    for (unsigned i = 0; i < nint;) {
        unsigned j = i;
        while (j < nint && S.bin_number[j] == S.bin_number[i])
            j++;
        S.bin_active.push_back(S.bin_number[i]);
        S.bin_start_index.push_back(i);
        i = j;
    }
    S.num_active_bins = (unsigned)S.bin_active.size();
    S.bin_start_index.push_back(nint);
    S.pair_shapeIDs.clear();
    if (S.num_active_bins == 0)
        return;
    const int nab = (int)S.num_active_bins;
    S.bin_num_contact.assign(nab + 1, 0);

    auto visit = [&](int index, long long* out) -> unsigned {  // f_Count/f_Store_AABB_AABB_Intersection :86-217
        unsigned start = S.bin_start_index[index], end = S.bin_start_index[index + 1];
        unsigned count = 0;
        if (end - start == 1)
            return 0;
        for (unsigned i = start; i < end; i++) {
            unsigned shapeA = S.bin_aabb_number[i];
            V3 Amin = S.aabb_min[shapeA], Amax = S.aabb_max[shapeA];
            unsigned bodyA = S.shapes[shapeA].body;
            if (S.collide[bodyA] == 0)
                continue;
            for (unsigned k = i + 1; k < end; k++) {
                unsigned shapeB = S.bin_aabb_number[k];
                unsigned bodyB = S.shapes[shapeB].body;
                if (shapeA == shapeB)
                    continue;
                if (bodyA == bodyB)
                    continue;
                if (S.collide[bodyB] == 0)
                    continue;
                if (!S.active[bodyA] && !S.active[bodyB])
                    continue;
                // family masks: all shapes use the defaults (1, 0x7FFF) -> collide() is true
                const V3& Bmin = S.aabb_min[shapeB];
                const V3& Bmax = S.aabb_max[shapeB];
                if (!overlap(Amin, Amax, Bmin, Bmax))
                    continue;
                if (!current_bin(Amin, Bmin, inv, bins, S.bin_active[index]))
                    continue;
                if (out) {
                    // reference swaps so that the smaller id goes in the high word (:201-206); with the
                    // stable sort above shapeA < shapeB always holds
                    unsigned a = std::min(shapeA, shapeB), b = std::max(shapeA, shapeB);
                    out[count] = ((long long)a << 32 | (long long)b);
                }
                count++;
            }
        }
        return count;
    };
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nab; i++)
        S.bin_num_contact[i] = visit(i, nullptr);
    {
        unsigned run = 0;
        for (int i = 0; i <= nab; i++) {
            unsigned c = S.bin_num_contact[i];
            S.bin_num_contact[i] = run;
            run += c;
        }
    }
    S.pair_shapeIDs.resize(S.bin_num_contact[nab]);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nab; i++)
        visit(i, S.pair_shapeIDs.data() + S.bin_num_contact[i]);
}

// ChNarrowphase::ProcessRigids, PRIMS dispatch restricted to sphere/box/triangle pairs
void narrowphase(System& S) {
    const int ns = (int)S.shapes.size();
    // PreprocessLocalToParent: ChNarrowphase.cpp:157-196
    S.shapeA_global.resize(ns);
    S.shapeR_global.resize(ns);
    S.tri_global.resize(3 * (size_t)ns);
#pragma omp parallel for
    for (int i = 0; i < ns; i++) {
        const Shape& sh = S.shapes[i];
        const V3& p = S.pos[sh.body];
        const Q4& r = S.rot[sh.body];
        S.shapeA_global[i] = TransformLocalToParent(p, r, sh.A);
        if (sh.type == TRIANGLE) {
            S.tri_global[3 * (size_t)i + 0] = TransformLocalToParent(p, r, sh.tri[0]);
            S.tri_global[3 * (size_t)i + 1] = TransformLocalToParent(p, r, sh.tri[1]);
            S.tri_global[3 * (size_t)i + 2] = TransformLocalToParent(p, r, sh.tri[2]);
        }
        S.shapeR_global[i] = Mult(r, sh.R);
    }
    const long long np = (long long)S.pair_shapeIDs.size();
    // scratch of the pair loop lives in the system: no per-step allocation / serial zero fill of ~90 B per pair
    std::vector<char>& act = S.np_act;
    std::vector<V3>&norm = S.np_norm, &p1 = S.np_p1, &p2 = S.np_p2;
    std::vector<double>&dep = S.np_dep, &er = S.np_er;
    if ((long long)act.size() < np) {
        act.resize(np); norm.resize(np); p1.resize(np); p2.resize(np); dep.resize(np); er.resize(np);
    }
    const double separation = 0;  // 2 * envelope, envelope = 0 (ChSystemMulticoreSMC.cpp:27)
#pragma omp parallel for
    for (long long idx = 0; idx < np; idx++) {  // DispatchPRIMS :267-288 + PRIMSCollision :1406-1531
        int a = int(S.pair_shapeIDs[idx] >> 32), b = int(S.pair_shapeIDs[idx] & 0xffffffff);
        const Shape& A = S.shapes[a];
        const Shape& B = S.shapes[b];
        bool hit = false;
        if (A.type == SPHERE && B.type == SPHERE) {
            hit = sphere_sphere(S.shapeA_global[a], A.dims.x, S.shapeA_global[b], B.dims.x, separation, norm[idx],
                                dep[idx], p1[idx], p2[idx], er[idx]);
        } else if (A.type == BOX && B.type == SPHERE) {
            hit = box_sphere(S.shapeA_global[a], S.shapeR_global[a], A.dims, S.shapeA_global[b], B.dims.x, separation,
                             norm[idx], dep[idx], p1[idx], p2[idx], er[idx]);
        } else if (A.type == SPHERE && B.type == BOX) {
            hit = box_sphere(S.shapeA_global[b], S.shapeR_global[b], B.dims, S.shapeA_global[a], A.dims.x, separation,
                             norm[idx], dep[idx], p2[idx], p1[idx], er[idx]);
            if (hit)
                norm[idx] = -norm[idx];
        } else if (A.type == TRIANGLE && B.type == SPHERE) {
            const V3* T = &S.tri_global[3 * (size_t)a];
            hit = triangle_sphere(T[0], T[1], T[2], S.shapeA_global[b], B.dims.x, separation, norm[idx], dep[idx],
                                  p1[idx], p2[idx], er[idx]);
        } else if (A.type == SPHERE && B.type == TRIANGLE) {
            const V3* T = &S.tri_global[3 * (size_t)b];
            hit = triangle_sphere(T[0], T[1], T[2], S.shapeA_global[a], A.dims.x, separation, norm[idx], dep[idx],
                                  p2[idx], p1[idx], er[idx]);
            if (hit)
                norm[idx] = -norm[idx];
        }
        act[idx] = hit;
    }
    // stream compaction (thrust::remove_if, order preserving): ChNarrowphase.cpp:362-385
    // exclusive scan of the flags, then a parallel scatter (same order as the serial compaction)
    std::vector<long long>& offs = S.np_offs;
    if ((long long)offs.size() < np + 1)
        offs.resize((size_t)np + 1);
    long long nkeep = 0;
    for (long long idx = 0; idx < np; idx++) {
        offs[idx] = nkeep;
        nkeep += act[idx] ? 1 : 0;
    }
    S.c_shape.resize(nkeep); S.c_b1.resize(nkeep); S.c_b2.resize(nkeep); S.c_norm.resize(nkeep); S.c_pt1.resize(nkeep);
    S.c_pt2.resize(nkeep); S.c_depth.resize(nkeep); S.c_erad.resize(nkeep);
#pragma omp parallel for
    for (long long idx = 0; idx < np; idx++) {
        if (!act[idx])
            continue;
        const long long o = offs[idx];
        int a = int(S.pair_shapeIDs[idx] >> 32), b = int(S.pair_shapeIDs[idx] & 0xffffffff);
        S.c_shape[o] = S.pair_shapeIDs[idx];
        S.c_b1[o] = S.shapes[a].body;
        S.c_b2[o] = S.shapes[b].body;
        S.c_norm[o] = norm[idx];
        S.c_pt1[o] = p1[idx];
        S.c_pt2[o] = p2[idx];
        S.c_depth[o] = dep[idx];
        S.c_erad[o] = er[idx];
    }
}

// ProcessContacts: ChIterativeSolverMulticoreSMC.cpp:644-725
int process_contacts(System& S) {
    const long long nc = (long long)S.c_shape.size();
    const int nb = (int)S.mass.size();
    ForceIn P{S.st.force_model, S.st.adhesion_model, S.st.tangential_mode, S.st.use_mat_props != 0, S.st.char_vel,
              S.st.min_slip_vel, S.st.min_roll_vel, S.st.min_spin_vel, S.st.dt};
    S.c_force.resize(nc);
    S.c_tq1.resize(nc);
    S.c_tq2.resize(nc);
    std::vector<int>& slot_of = S.pc_slot_of;
    std::vector<char>& is_new = S.pc_is_new;
    slot_of.resize(nc);
    is_new.resize(nc);
#pragma omp parallel for
    for (long long i = 0; i < nc; i++) {
        S.c_force[i] = V3(0);
        S.c_tq1[i] = V3(0);
        S.c_tq2[i] = V3(0);
        slot_of[i] = -1;
        is_new[i] = 0;
    }
    int err = 0;
    const bool multi = (P.displ_mode == ORC_TANG_MULTISTEP);
    if (multi) {
        {
            const long long nh = (long long)S.hist.size();
#pragma omp parallel for
            for (long long i = 0; i < nh; i++)
                S.hist[i].touch = 0;  // :661-662
        }
        // serial slot assignment in contact order (reference: inside the parallel loop, :194-227).
        // Contacts with depth >= 0 return before touching the history (:96-104).
        // Every thread walks all contacts in order but only serves the owner bodies of its own range: the rows fill in
        // exactly the order of the serial loop, without locks.
#pragma omp parallel reduction(| : err)
        {
        const int nth = omp_get_num_threads(), tid = omp_get_thread_num();
        const int b_lo = (int)((long long)nb * tid / nth), b_hi = (int)((long long)nb * (tid + 1) / nth);
        for (long long i = 0; i < nc; i++) {
            if (S.c_depth[i] >= 0)
                continue;
            int b1 = S.c_b1[i], b2 = S.c_b2[i];
            int sb1 = std::max(b1, b2), sb2 = std::min(b1, b2);
            if (sb1 < b_lo || sb1 >= b_hi)
                continue;
            int s1 = int(S.c_shape[i] >> 32), s2 = int(S.c_shape[i] & 0xffffffff);
            int ss1 = std::max(s1, s2), ss2 = std::min(s1, s2);
            HistSlot* row = &S.hist[(size_t)kMaxShear * sb1];
            int id = -1;
            for (int k = 0; k < kMaxShear; k++)
                if (row[k].nb == sb2 && row[k].s1 == ss1 && row[k].s2 == ss2) {
                    id = k;
                    break;
                }
            if (id < 0) {
                for (int k = 0; k < kMaxShear; k++)
                    if (row[k].nb == -1) {
                        id = k;
                        row[k].nb = sb2;
                        row[k].s1 = ss1;
                        row[k].s2 = ss2;
                        is_new[i] = 1;
                        break;
                    }
                if (id < 0) {
                    err = 1;  // reference would index slot -1 (out of bounds); report instead
                    continue;
                }
            }
            slot_of[i] = kMaxShear * sb1 + id;
        }
        }  // omp parallel
        if (err)
            return err;
    }
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < nc; i++) {
        int s1 = int(S.c_shape[i] >> 32), s2 = int(S.c_shape[i] & 0xffffffff);
        // AddContact(index,...): composite from the two shapes' materials
        Composite cm = make_composite(S.mats[S.shapes[s1].material], S.mats[S.shapes[s2].material]);
        HistSlot* slot = (multi && slot_of[i] >= 0) ? &S.hist[slot_of[i]] : nullptr;
        calc_contact_force(P, S.c_b1[i], S.c_b2[i], S.mass.data(), S.pos.data(), S.rot.data(), S.v.data(), cm,
                           S.c_pt1[i], S.c_pt2[i], S.c_norm[i], S.c_depth[i], S.c_erad[i], slot, is_new[i] != 0,
                           S.c_force[i], S.c_tq1[i], S.c_tq2[i]);
    }
    if (multi) {  // :677-685
        const long long nh = (long long)S.hist.size();
#pragma omp parallel for
        for (long long i = 0; i < nh; i++)
            if (!S.hist[i].touch)
                S.hist[i].nb = -1;
    }
    // reduce per body (:692-711); body 1 of a contact gets -force, body 2 gets +force
    S.body_force.resize(nb);
    S.body_torque.resize(nb);
    // same trick: per body the contributions are added in contact order, as the serial loop (and the reference's
    // sort-by-body + segmented sum) would
#pragma omp parallel
    {
        const int nth = omp_get_num_threads(), tid = omp_get_thread_num();
        const int b_lo = (int)((long long)nb * tid / nth), b_hi = (int)((long long)nb * (tid + 1) / nth);
        for (int b = b_lo; b < b_hi; b++) {
            S.body_force[b] = V3(0);
            S.body_torque[b] = V3(0);
        }
        for (long long i = 0; i < nc; i++) {
            const int b1 = S.c_b1[i], b2 = S.c_b2[i];
            if (b1 >= b_lo && b1 < b_hi) {
                S.body_force[b1] = S.body_force[b1] + (-S.c_force[i]);
                S.body_torque[b1] = S.body_torque[b1] + S.c_tq1[i];
            }
            if (b2 >= b_lo && b2 < b_hi) {
                S.body_force[b2] = S.body_force[b2] + S.c_force[i];
                S.body_torque[b2] = S.body_torque[b2] + S.c_tq2[i];
            }
        }
    }
    return 0;
}

void update_forces(System& S) {  // Update(): hf = h * Xforce, h * (Xtorque - gyro)
    const int nb = (int)S.mass.size();
    S.hf.assign((size_t)6 * nb, 0.0);
    const V3 g(S.st.gravity[0], S.st.gravity[1], S.st.gravity[2]);
    const double h = S.st.dt;
#pragma omp parallel for
    for (int i = 0; i < nb; i++) {
        V3 w(S.v[6 * i + 3], S.v[6 * i + 4], S.v[6 * i + 5]);
        S.gyro[i] = Cross(w, S.inertia[i] * w);  // ChBody.cpp:423-426
        V3 Xforce = g * S.mass[i];               // ChBody.cpp:554
        V3 Xtorque(0);
        V3 f = h * Xforce;                       // ChBody.cpp:247-256
        V3 t = h * (Xtorque - S.gyro[i]);
        S.hf[6 * i + 0] = f.x; S.hf[6 * i + 1] = f.y; S.hf[6 * i + 2] = f.z;
        S.hf[6 * i + 3] = t.x; S.hf[6 * i + 4] = t.y; S.hf[6 * i + 5] = t.z;
    }
}

int eval(System& S) {
    double t0 = now();
    update_forces(S);
    generate_aabb(S);
    broadphase(S);
    double t1 = now();
    narrowphase(S);
    double t2 = now();
    int err = process_contacts(S);
    double t3 = now();
    S.timers[0] += t1 - t0;
    S.timers[1] += t2 - t1;
    S.timers[2] += t3 - t2;
    return err;
}

int step(System& S) {
    double ts = now();
    int err = eval(S);
    if (err)
        return err;
    double t0 = now();
    const int nb = (int)S.mass.size();
    const double h = S.st.dt;
#pragma omp parallel for
    for (int i = 0; i < nb; i++) {
        if (!S.active[i])
            continue;
        // host_AddContactForces :604-619
        V3 cf = h * S.body_force[i];
        V3 ct = h * S.body_torque[i];
        double* hf = &S.hf[6 * (size_t)i];
        hf[0] += cf.x; hf[1] += cf.y; hf[2] += cf.z;
        hf[3] += ct.x; hf[4] += ct.y; hf[5] += ct.z;
        // M_invk = v + M_inv * hf; v = M_invk (ChIterativeSolverMulticore.cpp:158, SMC.cpp:839-853)
        double inv_mass = 1.0 / S.mass[i];
        double* v = &S.v[6 * (size_t)i];
        v[0] = v[0] + inv_mass * hf[0];
        v[1] = v[1] + inv_mass * hf[1];
        v[2] = v[2] + inv_mass * hf[2];
        v[3] = v[3] + S.inv_inertia[i].x * hf[3];
        v[4] = v[4] + S.inv_inertia[i].y * hf[4];
        v[5] = v[5] + S.inv_inertia[i].z * hf[5];
        // VariablesQbIncrementPosition: ChBody.cpp:288-310
        V3 newspeed(v[0], v[1], v[2]);
        V3 newwel(v[3], v[4], v[5]);
        S.pos[i] = S.pos[i] + newspeed * h;
        V3 newwel_abs = Rotate(newwel, S.rot[i]);  // GetRotMat() * newwel
        double len = Length(newwel_abs);
        double mangle = len * h;
        if (len < std::numeric_limits<double>::min())
            newwel_abs = V3(1, 0, 0);  // ChVector3::Normalize, ChVector3.h:873-883
        else
            newwel_abs = newwel_abs * (1 / len);
        double halfang = mangle / 2;
        double sinhalf = std::sin(halfang);
        Q4 dq(std::cos(halfang), newwel_abs.x * sinhalf, newwel_abs.y * sinhalf, newwel_abs.z * sinhalf);
        S.rot[i] = Mult(dq, S.rot[i]);
    }
    double t1 = now();
    S.timers[3] += t1 - t0;
    S.timers[4] += t1 - ts;
    return 0;
}

void c3(const double* p, V3& v) { v = V3(p[0], p[1], p[2]); }
void o3(const V3& v, double* p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

void* orc_create(const OrcSettings* s) {
    System* S = new System();
    S->st = *s;
    return S;
}
void orc_destroy(void* h) { delete (System*)h; }
void orc_set_settings(void* h, const OrcSettings* s) { ((System*)h)->st = *s; }

int orc_add_material(void* h, const OrcMaterial* m) {
    System& S = *(System*)h;
    S.mats.push_back(*m);
    return (int)S.mats.size() - 1;
}

int orc_add_body(void* h, double mass, const double inertia[3], const double pos[3], const double rot[4],
                 const double vel[3], const double omega_loc[3], int fixed) {
    System& S = *(System*)h;
    S.mass.push_back(mass);
    S.inertia.push_back(V3(inertia[0], inertia[1], inertia[2]));
    S.inv_inertia.push_back(V3(1.0 / inertia[0], 1.0 / inertia[1], 1.0 / inertia[2]));
    S.pos.push_back(V3(pos[0], pos[1], pos[2]));
    S.rot.push_back(Q4(rot[0], rot[1], rot[2], rot[3]));
    for (int k = 0; k < 3; k++)
        S.v.push_back(vel[k]);
    for (int k = 0; k < 3; k++)
        S.v.push_back(omega_loc[k]);
    S.active.push_back(fixed ? 0 : 1);
    S.collide.push_back(1);
    S.gyro.push_back(V3(0));
    for (int k = 0; k < kMaxShear; k++)  // ChSystemMulticoreSMC.cpp:41-49
        S.hist.push_back(HistSlot{-1, -1, -1, V3(0), 0.0, 0.0, 0});
    return (int)S.mass.size() - 1;
}

int orc_add_spheres(void* h, int n, const double* pos3, const double* vel3, const double* omega3,
                    const double* radius, const double* mass, int material) {
    System& S = *(System*)h;
    int first = (int)S.mass.size();
    const double zero[3] = {0, 0, 0};
    const double qid[4] = {1, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        double I = 0.4 * mass[i] * radius[i] * radius[i];
        double in3[3] = {I, I, I};
        int b = orc_add_body(h, mass[i], in3, pos3 + 3 * i, qid, vel3 ? vel3 + 3 * i : zero,
                             omega3 ? omega3 + 3 * i : zero, 0);
        Shape sh;
        sh.type = SPHERE;
        sh.body = b;
        sh.material = material;
        sh.A = V3(0);
        sh.R = Q4(1, 0, 0, 0);
        sh.dims = V3(radius[i], 0, 0);
        S.shapes.push_back(sh);
    }
    return first;
}

int orc_add_box(void* h, int body, int material, const double lpos[3], const double lrot[4], const double hdims[3]) {
    System& S = *(System*)h;
    Shape sh;
    sh.type = BOX;
    sh.body = body;
    sh.material = material;
    sh.A = V3(lpos[0], lpos[1], lpos[2]);
    sh.R = Q4(lrot[0], lrot[1], lrot[2], lrot[3]);
    sh.dims = V3(hdims[0], hdims[1], hdims[2]);
    S.shapes.push_back(sh);
    return (int)S.shapes.size() - 1;
}

int orc_add_triangles(void* h, int body, int material, int n, const double* v) {
    System& S = *(System*)h;
    int first = (int)S.shapes.size();
    for (int i = 0; i < n; i++) {
        Shape sh;
        sh.type = TRIANGLE;
        sh.body = body;
        sh.material = material;
        sh.A = V3(0);
        sh.R = Q4(1, 0, 0, 0);
        sh.dims = V3(0);
        for (int k = 0; k < 3; k++)
            sh.tri[k] = V3(v[9 * i + 3 * k], v[9 * i + 3 * k + 1], v[9 * i + 3 * k + 2]);
        S.shapes.push_back(sh);
    }
    return first;
}

void orc_set_body_state(void* h, int b, const double pos[3], const double rot[4], const double vel[3],
                        const double omega_loc[3]) {
    System& S = *(System*)h;
    if (pos) S.pos[b] = V3(pos[0], pos[1], pos[2]);
    if (rot) S.rot[b] = Q4(rot[0], rot[1], rot[2], rot[3]);
    if (vel) for (int k = 0; k < 3; k++) S.v[6 * (size_t)b + k] = vel[k];
    if (omega_loc) for (int k = 0; k < 3; k++) S.v[6 * (size_t)b + 3 + k] = omega_loc[k];
}
void orc_set_body_fixed(void* h, int b, int fixed) { ((System*)h)->active[b] = fixed ? 0 : 1; }

static void set_threads(System& S) {
#ifdef _OPENMP
    if (S.st.num_threads > 0)
        omp_set_num_threads(S.st.num_threads);
#endif
}

int orc_step(void* h, int nsteps) {
    System& S = *(System*)h;
    set_threads(S);
    for (int i = 0; i < nsteps; i++) {
        int e = step(S);
        if (e)
            return e;
    }
    return 0;
}
int orc_eval(void* h) {
    System& S = *(System*)h;
    set_threads(S);
    return eval(S);
}

int orc_num_bodies(void* h) { return (int)((System*)h)->mass.size(); }
int orc_num_shapes(void* h) { return (int)((System*)h)->shapes.size(); }

void orc_get_body_state(void* h, double* pos3, double* rot4, double* vel3, double* omega3) {
    System& S = *(System*)h;
    const size_t nb = S.mass.size();
    for (size_t i = 0; i < nb; i++) {
        if (pos3) o3(S.pos[i], pos3 + 3 * i);
        if (rot4) { rot4[4 * i] = S.rot[i].w; rot4[4 * i + 1] = S.rot[i].x; rot4[4 * i + 2] = S.rot[i].y; rot4[4 * i + 3] = S.rot[i].z; }
        if (vel3) for (int k = 0; k < 3; k++) vel3[3 * i + k] = S.v[6 * i + k];
        if (omega3) for (int k = 0; k < 3; k++) omega3[3 * i + k] = S.v[6 * i + 3 + k];
    }
}

void orc_get_grid(void* h, double origin[3], double bin_size[3], double inv_bin_size[3], int bins[3]) {
    System& S = *(System*)h;
    o3(S.origin, origin);
    o3(S.bin_size, bin_size);
    o3(S.inv_bin_size, inv_bin_size);
    for (int k = 0; k < 3; k++) bins[k] = S.st.bins_per_axis[k];
}
void orc_get_shape_bins(void* h, int* gmin3, int* gmax3) {
    System& S = *(System*)h;
    for (size_t i = 0; i < S.gmin.size(); i++) {
        gmin3[3 * i] = S.gmin[i].x; gmin3[3 * i + 1] = S.gmin[i].y; gmin3[3 * i + 2] = S.gmin[i].z;
        gmax3[3 * i] = S.gmax[i].x; gmax3[3 * i + 1] = S.gmax[i].y; gmax3[3 * i + 2] = S.gmax[i].z;
    }
}
void orc_get_broadphase_sizes(void* h, long long sizes[3]) {
    System& S = *(System*)h;
    sizes[0] = S.num_active_bins;
    sizes[1] = (long long)S.bin_number.size();
    sizes[2] = (long long)S.pair_shapeIDs.size();
}
void orc_get_bin_csr(void* h, unsigned* bin_active, unsigned* bin_start_index, unsigned* bin_aabb_number) {
    System& S = *(System*)h;
    if (bin_active) std::memcpy(bin_active, S.bin_active.data(), S.bin_active.size() * sizeof(unsigned));
    if (bin_start_index) std::memcpy(bin_start_index, S.bin_start_index.data(), S.bin_start_index.size() * sizeof(unsigned));
    if (bin_aabb_number) std::memcpy(bin_aabb_number, S.bin_aabb_number.data(), S.bin_aabb_number.size() * sizeof(unsigned));
}
void orc_get_pairs(void* h, long long* p) {
    System& S = *(System*)h;
    std::memcpy(p, S.pair_shapeIDs.data(), S.pair_shapeIDs.size() * sizeof(long long));
}
void orc_generate_aabb(void* h, double* min3, double* max3) {
    System& S = *(System*)h;
    generate_aabb(S);
    for (size_t i = 0; i < S.aabb_min.size(); i++) {
        o3(S.aabb_min[i], min3 + 3 * i);
        o3(S.aabb_max[i], max3 + 3 * i);
    }
}
long long orc_num_contacts(void* h) { return (long long)((System*)h)->c_shape.size(); }
void orc_get_contacts(void* h, long long* shape_pair, int* body_pair2, double* normal3, double* depth, double* pt1,
                      double* pt2, double* erad) {
    System& S = *(System*)h;
    for (size_t i = 0; i < S.c_shape.size(); i++) {
        if (shape_pair) shape_pair[i] = S.c_shape[i];
        if (body_pair2) { body_pair2[2 * i] = S.c_b1[i]; body_pair2[2 * i + 1] = S.c_b2[i]; }
        if (normal3) o3(S.c_norm[i], normal3 + 3 * i);
        if (depth) depth[i] = S.c_depth[i];
        if (pt1) o3(S.c_pt1[i], pt1 + 3 * i);
        if (pt2) o3(S.c_pt2[i], pt2 + 3 * i);
        if (erad) erad[i] = S.c_erad[i];
    }
}
void orc_get_contact_forces(void* h, double* f, double* t1, double* t2) {
    System& S = *(System*)h;
    for (size_t i = 0; i < S.c_force.size(); i++) {
        if (f) o3(S.c_force[i], f + 3 * i);
        if (t1) o3(S.c_tq1[i], t1 + 3 * i);
        if (t2) o3(S.c_tq2[i], t2 + 3 * i);
    }
}
void orc_get_body_forces(void* h, double* force3, double* torque3) {
    System& S = *(System*)h;
    for (size_t i = 0; i < S.body_force.size(); i++) {
        if (force3) o3(S.body_force[i], force3 + 3 * i);
        if (torque3) o3(S.body_torque[i], torque3 + 3 * i);
    }
}
long long orc_num_history(void* h) {
    System& S = *(System*)h;
    long long n = 0;
    for (auto& s : S.hist) n += (s.nb != -1);
    return n;
}
void orc_get_history(void* h, int* body, int* other, int* shape1, int* shape2, double* disp3, double* duration,
                     double* relvel_init) {
    System& S = *(System*)h;
    size_t n = 0;
    for (size_t i = 0; i < S.hist.size(); i++) {
        const HistSlot& s = S.hist[i];
        if (s.nb == -1) continue;
        body[n] = (int)(i / kMaxShear);
        other[n] = s.nb; shape1[n] = s.s1; shape2[n] = s.s2;
        o3(s.disp, disp3 + 3 * n);
        duration[n] = s.duration;
        relvel_init[n] = s.relvel_init;
        n++;
    }
}
void orc_add_history(void* h, int body, int other, int shape1, int shape2, const double disp[3], double duration,
                     double relvel_init) {
    System& S = *(System*)h;
    HistSlot* row = &S.hist[(size_t)kMaxShear * body];
    for (int k = 0; k < kMaxShear; k++)
        if (row[k].nb == -1) {
            row[k] = HistSlot{other, shape1, shape2, V3(disp[0], disp[1], disp[2]), relvel_init, duration, 0};
            return;
        }
}
void orc_get_timers(void* h, double t[5]) { std::memcpy(t, ((System*)h)->timers, sizeof(double) * 5); }
void orc_reset_timers(void* h) { std::memset(((System*)h)->timers, 0, sizeof(double) * 5); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_sphere_sphere(const double pos1[3], double r1, const double pos2[3], double r2, double separation,
                      double norm[3], double* depth, double pt1[3], double pt2[3], double* erad) {
    V3 a, b, n, p1, p2;
    c3(pos1, a); c3(pos2, b);
    double d = 0, e = 0;
    bool hit = sphere_sphere(a, r1, b, r2, separation, n, d, p1, p2, e);
    if (hit) { o3(n, norm); o3(p1, pt1); o3(p2, pt2); *depth = d; *erad = e; }
    return hit;
}
int orc_box_sphere(const double pos1[3], const double rot1[4], const double hdims1[3], const double pos2[3],
                   double r2, double separation, double norm[3], double* depth, double pt1[3], double pt2[3],
                   double* erad) {
    V3 a, hd, b, n, p1, p2;
    c3(pos1, a); c3(hdims1, hd); c3(pos2, b);
    Q4 q(rot1[0], rot1[1], rot1[2], rot1[3]);
    double d = 0, e = 0;
    bool hit = box_sphere(a, q, hd, b, r2, separation, n, d, p1, p2, e);
    if (hit) { o3(n, norm); o3(p1, pt1); o3(p2, pt2); *depth = d; *erad = e; }
    return hit;
}
int orc_triangle_sphere(const double A[3], const double B[3], const double C[3], const double pos2[3], double r2,
                        double separation, double norm[3], double* depth, double pt1[3], double pt2[3],
                        double* erad) {
    V3 a, b, c, p, n, p1, p2;
    c3(A, a); c3(B, b); c3(C, c); c3(pos2, p);
    double d = 0, e = 0;
    bool hit = triangle_sphere(a, b, c, p, r2, separation, n, d, p1, p2, e);
    if (hit) { o3(n, norm); o3(p1, pt1); o3(p2, pt2); *depth = d; *erad = e; }
    return hit;
}
unsigned orc_snap_to_box(const double hdims[3], double loc[3]) {
    V3 hd, l;
    c3(hdims, hd); c3(loc, l);
    unsigned code = snap_to_box(hd, l);
    o3(l, loc);
    return code;
}
int orc_snap_to_triangle(const double A[3], const double B[3], const double C[3], const double P[3], double res[3]) {
    V3 a, b, c, p, r;
    c3(A, a); c3(B, b); c3(C, c); c3(P, p);
    bool edge = snap_to_triangle(a, b, c, p, r);
    o3(r, res);
    return edge;
}
void orc_rotate(const double v[3], const double q[4], double out[3], double outT[3], double outAbs[3]) {
    V3 a;
    c3(v, a);
    Q4 Q(q[0], q[1], q[2], q[3]);
    o3(Rotate(a, Q), out);
    o3(RotateT(a, Q), outT);
    o3(AbsRotate(Q, a), outAbs);
}
void orc_composite(const OrcMaterial* m1, const OrcMaterial* m2, double out[13]) {
    Composite c = make_composite(*m1, *m2);
    const float* f = &c.E_eff;
    for (int i = 0; i < 13; i++) out[i] = f[i];
}

int orc_contact_force(const OrcSettings* s, const double comp[13], int b1, int b2, const double* mass,
                      const double* pos, const double* rot, const double* vel, const double pt1[3],
                      const double pt2[3], const double normal[3], double depth, double erad, int* hist_present,
                      double hist_disp[3], double* hist_dur, double* hist_relvel, double force_b2[3],
                      double torque_b1[3], double torque_b2[3]) {
    ForceIn P{s->force_model, s->adhesion_model, s->tangential_mode, s->use_mat_props != 0, s->char_vel,
              s->min_slip_vel, s->min_roll_vel, s->min_spin_vel, s->dt};
    Composite cm;
    float* f = &cm.E_eff;
    for (int i = 0; i < 13; i++) f[i] = (float)comp[i];
    V3 p[2] = {V3(pos[0], pos[1], pos[2]), V3(pos[3], pos[4], pos[5])};
    Q4 q[2] = {Q4(rot[0], rot[1], rot[2], rot[3]), Q4(rot[4], rot[5], rot[6], rot[7])};
    HistSlot slot{-1, -1, -1, V3(0), 0.0, 0.0, 0};
    bool newc = true;
    if (*hist_present) {
        slot.nb = std::min(b1, b2);
        slot.disp = V3(hist_disp[0], hist_disp[1], hist_disp[2]);
        slot.duration = *hist_dur;
        slot.relvel_init = *hist_relvel;
        newc = false;
    }
    V3 P1, P2, N, F, T1, T2;
    c3(pt1, P1); c3(pt2, P2); c3(normal, N);
    calc_contact_force(P, b1, b2, mass, p, q, vel, cm, P1, P2, N, depth, erad, &slot, newc, F, T1, T2);
    if (s->tangential_mode == ORC_TANG_MULTISTEP && depth < 0) {
        *hist_present = 1;
        o3(slot.disp, hist_disp);
        *hist_dur = slot.duration;
        *hist_relvel = slot.relvel_init;
    }
    o3(F, force_b2); o3(T1, torque_b1); o3(T2, torque_b2);
    return 0;
}

}  // extern "C"

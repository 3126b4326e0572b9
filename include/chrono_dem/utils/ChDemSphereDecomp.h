// Stand-in for chrono_dem/utils/ChDemSphereDecomp.h (reference: src/chrono_dem/utils/ChDemSphereDecomp.h:33-125):
// approximate a triangle mesh by equal spheres whose centres lie in the facets -- used by demo_DEM_fixedTerrain to build
// a terrain of fixed particles.  Same signature; own construction: every facet is covered by rows of spheres parallel to
// the edge opposite its smallest angle, rows 2 r apart and inset by r from the edges, spheres at most 2 r apart within
// a row, and a last file of spheres running into the acute corner.
#ifndef CHRONO_B200_CHDEMSPHEREDECOMP_H
#define CHRONO_B200_CHDEMSPHEREDECOMP_H
#include <cmath>
#include <string>
#include <vector>

#include "chrono/core/ChVector3.h"
#include "chrono/geometry/ChTriangleMeshConnected.h"

namespace chrono {
namespace dem {

template <typename Real>
std::vector<ChVector3<Real>> MeshSphericalDecomposition(std::string objfilename,  ///< OBJ mesh file path
                                                        ChVector3<Real> scaling,  ///< scaling applied to the mesh first
                                                        ChVector3<Real> offset,   ///< displacement applied after scaling
                                                        Real sphere_radius        ///< radius of all spheres
) {
    std::vector<ChVector3<Real>> out;
    ChTriangleMeshConnected mesh;
    if (!mesh.LoadWavefrontMesh(objfilename, true, false))
        return out;
    mesh.Transform(ChVector3d(offset), ChMatrix33<double>(ChVector3d(scaling)));
    const double r = (double)sphere_radius;
    for (unsigned int f = 0; f < mesh.GetNumTriangles(); f++) {
        const ChTriangle t = mesh.GetTriangle(f);
        ChVector3d V[3] = {t.p1, t.p2, t.p3};
        // corner with the smallest angle becomes A; the rows run parallel to the opposite edge BC
        int ia = 0;
        double best = 1e300;
        for (int k = 0; k < 3; k++) {
            const ChVector3d u = (V[(k + 1) % 3] - V[k]).GetNormalized(), w = (V[(k + 2) % 3] - V[k]).GetNormalized();
            const double ang = std::acos(std::max(-1.0, std::min(1.0, u.Dot(w))));
            if (ang < best) { best = ang; ia = k; }
        }
        const ChVector3d A = V[ia], B = V[(ia + 1) % 3], C = V[(ia + 2) % 3];
        const ChVector3d bc = (C - B).GetNormalized();
        ChVector3d h = (A - B) - bc * (A - B).Dot(bc);  // from the line BC to A, in the plane
        const double H = h.Length();
        if (!(H > 0) || !((C - B).Length() > 0))
            continue;
        h = h / H;
        bool any = false;
        for (double d = r; d < H; d += 2 * r) {
            // the row at distance d from BC spans between the two other edges, inset by r
            const double s = 1.0 - d / H;  // similar triangle
            const ChVector3d P = A + (B - A) * s, Q = A + (C - A) * s;
            const double L = (Q - P).Length() - 2 * r;
            if (L <= 0) {
                out.push_back(ChVector3<Real>((P + Q) * 0.5));
                any = true;
                continue;
            }
            const int n = (int)std::ceil(L / (2 * r));
            const ChVector3d dir = (Q - P).GetNormalized();
            for (int j = 0; j <= n; j++)
                out.push_back(ChVector3<Real>(P + dir * (r + L * j / n)));
            any = true;
        }
        if (!any)  // a sliver thinner than a sphere: one sphere at its centroid
            out.push_back(ChVector3<Real>((A + B + C) / 3.0));
    }
    return out;
}

}  // namespace dem
}  // namespace chrono
#endif

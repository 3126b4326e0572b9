// Stand-in that shadows chrono/collision/ChCollisionModel.h for the standalone build of the
// reference's Multicore collision sources.  Provides only what those sources use:
//  - ChCollisionShape::Type, same enumerators in the same order as
//    src/chrono/collision/ChCollisionShape.h:31-53 (numeric identity matters for dispatch);
//  - ChCollisionInfo::GetDefaultEffectiveCurvatureRadius (src/chrono/collision/ChCollisionInfo.cpp:69,
//    default 0.1 at :24).
#pragma once
#include <vector>
#include <memory>
#include <climits>
#include "chrono/core/ChApiCE.h"
namespace chrono {
class ChCollisionShape {
  public:
    enum Type {
        SPHERE, ELLIPSOID, BOX, ROUNDEDBOX, CYLINDER, CYLSHELL, ROUNDEDCYL, CAPSULE, CONVEXHULL,
        TRIANGLEMESH, BARREL, POINT, SEGMENT, TRIANGLE, CONNECTEDTRIANGLE, CONE, TETRAHEDRON,
        PATH2D, SEGMENT2D, ARC2D, UNKNOWN_SHAPE
    };
};
class ChCollisionInfo {
  public:
    static double GetDefaultEffectiveCurvatureRadius() { return 0.1; }
};
}  // namespace chrono

// Stand-in that shadows chrono/physics/ChContactContainer.h when compiling the reference's
// Multicore collision sources standalone.  The real header drags in Chrono core (-> Eigen3, which
// is neither installed nor vendored); ChCollisionData.h only needs <vector> from it.
#pragma once
#include <vector>
#include <memory>
#include <climits>
#include "chrono/core/ChApiCE.h"

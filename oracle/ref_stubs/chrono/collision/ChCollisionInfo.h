// Stand-in that shadows chrono/collision/ChCollisionInfo.h; see ChCollisionModel.h stand-in.
#pragma once
#include "chrono/collision/ChCollisionModel.h"

// Force-included (-include) ahead of every reference translation unit in the standalone build:
// supplies the std headers and Chrono constants that the real headers would have reached through
// Chrono core (Eigen-dependent, not available here).
#pragma once
#include <vector>
#include <memory>
#include <climits>
#include <cmath>
#include <algorithm>
#include "chrono/utils/ChConstants.h"
#include "chrono/collision/ChCollisionModel.h"

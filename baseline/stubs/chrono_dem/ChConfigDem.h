// Stand-in for the CMake-generated chrono_dem/ChConfigDem.h (src/chrono_dem/ChConfigDem.h.in).
#pragma once
#include "chrono/ChConfig.h"

// Stand-in: src/chrono_dem/physics/ChSystemDem_impl.cpp includes chrono/core/ChVector3.h only for the constants it
// pulls in (CH_PI, CH_4_3); the real header needs Eigen3, which this image does not have.  The constants header is the
// reference's own.
#pragma once
#include "chrono/utils/ChConstants.h"

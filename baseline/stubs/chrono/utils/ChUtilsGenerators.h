// Empty stand-in: src/chrono_dem/physics/ChSystemDem_impl.cpp includes this header but uses nothing from it; the real one
// needs Eigen3, which this image does not have.
#pragma once

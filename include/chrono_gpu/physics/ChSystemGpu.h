// Old spelling of the module (BASELINE.json north star: ChSystemGpu / ChSystemGpuMesh): aliases of chrono::dem::*.
#ifndef CHRONO_B200_CHSYSTEMGPU_H
#define CHRONO_B200_CHSYSTEMGPU_H
#include "chrono_dem/physics/ChSystemDem.h"
#endif

// Stand-in for the CMake-generated chrono/ChVersion.h.
#pragma once
#define CHRONO_VERSION_MAJOR 10
#define CHRONO_VERSION_MINOR 0
#define CHRONO_VERSION_PATCH 0

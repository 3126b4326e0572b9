// Public enums / typedefs of Chrono::Dem, same names and values as the reference (src/chrono_dem/ChDemDefines.h:24-54).
#ifndef CHRONO_B200_CHDEMDEFINES_H
#define CHRONO_B200_CHDEMDEFINES_H
#include <climits>
#include <cstddef>
#include <functional>

#if __has_include(<vector_types.h>)
#include <vector_types.h>      // CUDA's double3 / float3, as in the reference
#include <vector_functions.h>  // make_double3
#else
struct double3 { double x, y, z; };
struct float3 { float x, y, z; };
inline double3 make_double3(double x, double y, double z) { return double3{x, y, z}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
#endif

namespace chrono {
namespace dem {

/// Used to compute position as a function of time (ChDemDefines.h:27).
typedef std::function<double3(float)> GranPositionFunction;
const GranPositionFunction GranPosFunction_default = [](float t) { return make_double3(0, 0, 0); };

enum class CHDEM_VERBOSITY { QUIET = 0, INFO = 1, METRICS = 2 };
enum class CHDEM_MESH_VERBOSITY { QUIET = 0, INFO = 1 };
enum class CHDEM_OUTPUT_MODE { CSV, BINARY, HDF5, NONE };
enum class CHDEM_TIME_INTEGRATOR { FORWARD_EULER, CHUNG, CENTERED_DIFFERENCE, EXTENDED_TAYLOR };
enum class CHDEM_FRICTION_MODE { FRICTIONLESS, SINGLE_STEP, MULTI_STEP };
enum class CHDEM_ROLLING_MODE { NO_RESISTANCE, SCHWARTZ, ELASTIC_PLASTIC };
enum CHDEM_RUN_MODE { FRICTIONLESS = 0, ONE_STEP = 1, MULTI_STEP = 2 };
enum CHDEM_OUTPUT_FLAGS { ABSV = 1, VEL_COMPONENTS = 2, FIXITY = 4, ANG_VEL_COMPONENTS = 8, FORCE_COMPONENTS = 16 };

}  // namespace dem
}  // namespace chrono

constexpr size_t BD_WALL_ID_X_BOT = 0;
constexpr size_t BD_WALL_ID_X_TOP = 1;
constexpr size_t BD_WALL_ID_Y_BOT = 2;
constexpr size_t BD_WALL_ID_Y_TOP = 3;
constexpr size_t BD_WALL_ID_Z_BOT = 4;
constexpr size_t BD_WALL_ID_Z_TOP = 5;
constexpr size_t NUM_RESERVED_BC_IDS = 6;
#define MAX_SPHERES_TOUCHED_BY_SPHERE 12
#define NULL_CHDEM_ID UINT_MAX

#endif

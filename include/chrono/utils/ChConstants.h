// Minimal stand-in for chrono/utils/ChConstants.h.
#ifndef CHRONO_B200_CHCONSTANTS_H
#define CHRONO_B200_CHCONSTANTS_H
namespace chrono {
static constexpr double CH_PI = 3.141592653589793238462643383279;
static constexpr double CH_PI_2 = 1.570796326794896619231321691639;
static constexpr double CH_PI_3 = 1.047197551196597746154214461093;
static constexpr double CH_PI_4 = 0.785398163397448309615660845819;
static constexpr double CH_2PI = 6.283185307179586476925286766559;
static constexpr double CH_DEG_TO_RAD = CH_PI / 180.0;
static constexpr double CH_RAD_TO_DEG = 180.0 / CH_PI;
}  // namespace chrono
#endif

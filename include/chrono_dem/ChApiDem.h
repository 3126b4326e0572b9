// Stand-in for src/chrono_dem/ChApiDem.h: export macro of the Chrono::Dem module.
#ifndef CHRONO_B200_CHAPIDEM_H
#define CHRONO_B200_CHAPIDEM_H
#if defined(_WIN32)
#define CH_DEM_API __declspec(dllexport)
#else
#define CH_DEM_API __attribute__((visibility("default")))
#endif
#endif

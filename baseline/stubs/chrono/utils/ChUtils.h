// Stand-in: src/chrono_dem/physics/ChSystemDem_impl.cpp uses nothing from chrono/utils/ChUtils.h itself, only the standard
// headers it drags in; the real header needs Eigen3, which this image does not have.
#pragma once
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <cstdio>
#include <stdexcept>
// same contract as the macro of the real header (chrono/utils/ChUtils.h:42): throw std::runtime_error when the expression is false
#ifndef ChAssertAlways
#define ChAssertAlways(exp)                                                                                  \
    {                                                                                                        \
        if (!(exp)) {                                                                                        \
            char msg_[300];                                                                                  \
            std::snprintf(msg_, sizeof(msg_), "Expression '%s' returned false - file %s, line %d.", #exp, __FILE__, __LINE__); \
            throw std::runtime_error(msg_);                                                                  \
        }                                                                                                    \
    }
#endif

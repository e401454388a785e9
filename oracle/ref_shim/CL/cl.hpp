// Shim for <CL/cl.hpp>: only the scalar typedefs and standard headers that the reference's CPU test helpers
// (/root/reference/include/ICP/tests/helper_funcs.hpp) need.  No OpenCL is available in this image
// (SURVEY.md section 0, F3).  Written for this repo; not derived from the Khronos header.
#pragma once
#include <cstdint>
#include <cmath>
#include <cassert>
#include <iostream>
#include <iomanip>
#include <numeric>
#include <algorithm>
#include <functional>
#include <sys/types.h>
typedef float    cl_float;
typedef double   cl_double;
typedef uint32_t cl_uint;
typedef int32_t  cl_int;
typedef uint16_t cl_ushort;
typedef uint8_t  cl_uchar;

// oracle/ref_build/shim/ref_prefix.h — force-included (-include) ahead of the reference sources by oracle/ref_build/build.sh.
// glibc >= 2.27 with GCC >= 13 spells M_PIf32 as a _Float32 literal (3.14...f32); nvcc 12.9's device front end has no
// _Float32 ("Internal Compiler Error (codegen): unsupported float variant") and the reference's utils/angle_utils.cuh:22-25
// uses M_PIf32 in __host__ __device__ code. On the toolchains the reference was written for the macro is a plain float
// literal; restore that meaning. Test infrastructure: nothing here is part of the product.
#pragma once
#include <math.h>
#include <cmath>
#ifdef M_PIf32
#undef M_PIf32
#endif
#define M_PIf32 3.14159265358979323846f
#ifdef M_PI_2f32
#undef M_PI_2f32
#endif
#define M_PI_2f32 1.57079632679489661923f
#ifdef M_PI_4f32
#undef M_PI_4f32
#endif
#define M_PI_4f32 0.78539816339744830962f

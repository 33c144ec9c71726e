/*
 * strugepic_user_w.h -- the user-supplied interpolation slot of strugepic_b200.
 *
 * The reference declares its interpolation functions in include/strugepic_w.hpp:12-16
 * (W1, Wp, I_W1, I_Wp, interpolation_range) and defines the defaults as WEAK symbols
 * (src/interpolation/interpolation.cpp:10,14,20,89): "How and where they are implemented is up
 * to the user (just link against a compiled object file or dynamic library)".  On the GPU the
 * same freedom is a link step too: write ONE source file that defines the five symbols below,
 *
 *     #include "strugepic_user_w.h"
 *     SPIC_W_CONST int spic_user_interpolation_range = 2;          // 1 or 2
 *     SPIC_W_FN double spic_user_W1(double x)            { ... }   // support (-R, R)
 *     SPIC_W_FN double spic_user_Wp(double x)            { ... }   // support (-R+1, R)
 *     SPIC_W_FN double spic_user_I_W1(double a, double b){ ... }   // integral of W1 over [a,b]
 *     SPIC_W_FN double spic_user_I_Wp(double a, double b){ ... }   // integral of Wp over [a,b]
 *
 * and build the library with it:  python -m strugepic_b200.build --user-w my_w.cu --out libmine.so
 * nvcc compiles the file as relocatable device code (__host__ __device__) and device-links it with
 * the library's generic particle kernels (csrc/user_w.cu); contexts created with
 * interp = SPIC_INTERP_USER then run every sub-flow with these functions (thread-per-particle
 * engine; the tap-specialised warp-per-cell kernels exist for the two shipped variants only).
 * The stock library carries csrc/user_w_default.cu in the slot: the cubic B-spline pair.
 *
 * Charge conservation needs  W1'(x) = Wp(x+1) - Wp(x)  and partition of unity of W1 and Wp
 * (SURVEY.md section 8c).  The same file compiles as plain C++, which is how the test-suite
 * links it into the reference itself (oracle/build_oracle.py: build_ref_user).
 */
#ifndef STRUGEPIC_USER_W_H
#define STRUGEPIC_USER_W_H

#if defined(__CUDACC__)
#define SPIC_W_FN extern "C" __host__ __device__
#define SPIC_W_CONST extern "C" const
#elif defined(__cplusplus)
#define SPIC_W_FN extern "C"
#define SPIC_W_CONST extern "C" const
#else
#define SPIC_W_FN
#define SPIC_W_CONST const
#endif

SPIC_W_CONST int spic_user_interpolation_range;
SPIC_W_FN double spic_user_W1(double x);
SPIC_W_FN double spic_user_Wp(double x);
SPIC_W_FN double spic_user_I_W1(double a, double b);
SPIC_W_FN double spic_user_I_Wp(double a, double b);

#endif

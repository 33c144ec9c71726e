// Default content of the user-W slot (include/strugepic_user_w.h): the cubic B-spline pair.
//
//   W1(x) = B3(x)            cubic B-spline, support (-2, 2)            -> interpolation_range 2
//   Wp(x) = B2(x - 1/2)      quadratic B-spline shifted to (-1, 2)
// B3'(x) = B2(x + 1/2) - B2(x - 1/2) = Wp(x + 1) - Wp(x): the identity charge conservation needs
// (include/strugepic_w.hpp:12-16 of the reference).  Both are partitions of unity.
// Written as an example of what a user supplies: plain functions, no tables, valid as C++ and as CUDA.
#include "../../include/strugepic_user_w.h"

#include <math.h>

SPIC_W_CONST int spic_user_interpolation_range = 2;

SPIC_W_FN double spic_user_W1(double x) {
  const double a = fabs(x);
  if (a >= 2.0) return 0.0;
  if (a >= 1.0) {
    const double t = 2.0 - a;
    return t * t * t * (1.0 / 6.0);
  }
  return 2.0 / 3.0 - a * a + 0.5 * a * a * a;
}

// quadratic B-spline and its running integral, centred at 0
static
#if defined(__CUDACC__)
    __host__ __device__
#endif
    inline double b2(double t) {
  const double a = fabs(t);
  if (a >= 1.5) return 0.0;
  if (a >= 0.5) {
    const double u = 1.5 - a;
    return 0.5 * u * u;
  }
  return 0.75 - a * a;
}
static
#if defined(__CUDACC__)
    __host__ __device__
#endif
    inline double b2_cdf(double t) {
  if (t <= -1.5) return 0.0;
  if (t >= 1.5) return 1.0;
  if (t < -0.5) {
    const double u = t + 1.5;
    return u * u * u * (1.0 / 6.0);
  }
  if (t > 0.5) {
    const double u = 1.5 - t;
    return 1.0 - u * u * u * (1.0 / 6.0);
  }
  return 0.5 + 0.75 * t - t * t * t * (1.0 / 3.0);
}
static
#if defined(__CUDACC__)
    __host__ __device__
#endif
    inline double b3_cdf(double x) {
  if (x <= -2.0) return 0.0;
  if (x >= 2.0) return 1.0;
  if (x < -1.0) {
    const double u = x + 2.0;
    return u * u * u * u * (1.0 / 24.0);
  }
  if (x > 1.0) {
    const double u = 2.0 - x;
    return 1.0 - u * u * u * u * (1.0 / 24.0);
  }
  const double x3 = x * x * x;
  return 0.5 + (2.0 / 3.0) * x - x3 * (1.0 / 3.0) + (x < 0.0 ? -1.0 : 1.0) * x3 * x * 0.125;
}

SPIC_W_FN double spic_user_Wp(double x) { return b2(x - 0.5); }
SPIC_W_FN double spic_user_I_W1(double a, double b) { return b3_cdf(b) - b3_cdf(a); }
SPIC_W_FN double spic_user_I_Wp(double a, double b) { return b2_cdf(b - 0.5) - b2_cdf(a - 0.5); }

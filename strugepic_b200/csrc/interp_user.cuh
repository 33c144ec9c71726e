// The Interp concept of interp.cuh over interpolation functions that are only DECLARED
// (include/strugepic_user_w.h) and defined in another translation unit: the file the user hands to
// `python -m strugepic_b200.build --user-w`, or csrc/user_w_default.cu.  Every translation unit that instantiates a
// kernel with it is compiled with -rdc=true and device-linked with the user's definitions.
//
// Replaces the reference's link-time override of the weak W1 / Wp / I_W1 / I_Wp / interpolation_range
// (include/strugepic_w.hpp:12-16, src/interpolation/interpolation.cpp:10,14,20,89).
#pragma once
#include "../../include/strugepic_user_w.h"
#include "interp.cuh"

namespace spic {

template <int R>  // R = interpolation_range
struct InterpUser {
  static constexpr int W = R;
  static constexpr int NW1 = 2 * R;      // offsets -R+1 .. R      (propagators.hpp:83-85)
  static constexpr int NWP = 2 * R - 1;  // offsets -R+1 .. R-1    (propagators.hpp:86-88)
  static SPIC_HDI double W1(double x) { return spic_user_W1(x); }
  static SPIC_HDI double Wp(double x) { return spic_user_Wp(x); }
  static SPIC_HDI double I_W1(double a, double b) { return spic_user_I_W1(a, b); }
  static SPIC_HDI double I_Wp(double a, double b) { return spic_user_I_Wp(a, b); }
  template <int T>
  static SPIC_HDI double w1_tap(double a) { return spic_user_W1(a); }
  template <int T>
  static SPIC_HDI double wp_tap(double a) { return spic_user_Wp(a); }
  template <int T>
  static SPIC_HDI double iwp_seg(double a, double b) { return spic_user_I_Wp(a, b); }
  // in-cell forms of the warp-per-cell kernels: the same calls (a user function has no per-tap pieces to pick from)
  template <int T>
  static SPIC_HDI double w1_in(double a) { return spic_user_W1(a); }
  template <int T>
  static SPIC_HDI double wp_in(double a) { return spic_user_Wp(a); }
  template <int T>
  static SPIC_HDI double iwp_seg_in(double a, double b) { return spic_user_I_Wp(a, b); }
};

// (InterpUser<1> / <2> by the range the linked-in user file defines)
#define SPIC_USER_DISPATCH(call)              \
  do {                                        \
    if (spic_user_interpolation_range == 2) { \
      using I = InterpUser<2>;                \
      call;                                   \
    } else {                                  \
      using I = InterpUser<1>;                \
      call;                                   \
    }                                         \
  } while (0)

}  // namespace spic

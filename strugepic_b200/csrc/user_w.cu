// The user-W slot: the thread-per-particle sub-flow kernels instantiated for interpolation functions that are
// only DECLARED here (include/strugepic_user_w.h) and defined in another translation unit -- the file the user
// hands to `python -m strugepic_b200.build --user-w`, or csrc/user_w_default.cu.  This file and the user's are
// compiled with -rdc=true and device-linked; every other kernel of the library stays whole-program.
//
// Replaces the reference's link-time override of the weak W1 / Wp / I_W1 / I_Wp / interpolation_range
// (include/strugepic_w.hpp:12-16, src/interpolation/interpolation.cpp:10,14,20,89).
#include "../../include/strugepic_user_w.h"
#include "particles_direct.cuh"

namespace spic {

// the Interp concept of interp.cuh over external functions; R = interpolation_range
template <int R>
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
};

int user_w_range() { return spic_user_interpolation_range; }

#define SPIC_USER_DISPATCH(call)                          \
  do {                                                    \
    if (spic_user_interpolation_range == 2) {             \
      using I = InterpUser<2>;                            \
      call;                                               \
    } else {                                              \
      using I = InterpUser<1>;                            \
      call;                                               \
    }                                                     \
  } while (0)

void user_theta_axis(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double m,
                     int comp, double dt) {
  SPIC_USER_DISPATCH(direct::theta_axis_launch<I>(c, p, n, n_dev, q, m, comp, dt));
}
void user_push_v_e(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double coef) {
  SPIC_USER_DISPATCH(direct::push_v_e_launch<I>(c, p, n, n_dev, coef));
}
void user_deposit_rho(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double q, double* out) {
  SPIC_USER_DISPATCH(direct::deposit_rho_launch<I>(c, p, n, n_dev, q, out));
}
void user_number_density(Ctx* c, const ParticleSoA& p, long n, const unsigned long long* n_dev, double* nd) {
  SPIC_USER_DISPATCH(direct::number_density_launch<I>(c, p, n, n_dev, nd));
}

}  // namespace spic

// host evaluation for spic_W1(SPIC_INTERP_USER, x) etc. (api.cu)
extern "C" {
double spic_user_host_W1(double x) { return spic_user_W1(x); }
double spic_user_host_Wp(double x) { return spic_user_Wp(x); }
double spic_user_host_I_W1(double a, double b) { return spic_user_I_W1(a, b); }
double spic_user_host_I_Wp(double a, double b) { return spic_user_I_Wp(a, b); }
}

// The user-W slot: the thread-per-particle sub-flow kernels instantiated for interpolation functions that are
// only DECLARED here (include/strugepic_user_w.h) and defined in another translation unit -- the file the user
// hands to `python -m strugepic_b200.build --user-w`, or csrc/user_w_default.cu.  This file and the user's are
// compiled with -rdc=true and device-linked; every other kernel of the library stays whole-program.
//
// Replaces the reference's link-time override of the weak W1 / Wp / I_W1 / I_Wp / interpolation_range
// (include/strugepic_w.hpp:12-16, src/interpolation/interpolation.cpp:10,14,20,89).
#include "interp_user.cuh"
#include "particles_direct.cuh"

namespace spic {

int user_w_range() { return spic_user_interpolation_range; }

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

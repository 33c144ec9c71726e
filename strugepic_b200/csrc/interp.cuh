// Interpolation (Whitney-form) functions W1, Wp, I_W1, I_Wp for the device.
//
// Interface: include/strugepic_w.hpp:12-16 of the reference; default definitions
// src/interpolation/interpolation.cpp (P8R2: 19-84, PWL: 87-157), evaluator
// src/interpolation/poly_util.hpp:19-51.  Both variants are compiled into every
// particle kernel as a template parameter (the reference picks one at link time
// through weak symbols).
//
// Tap form.  Every call site in the sub-flows evaluates W at `x - (cell + o)` for
// a fixed integer offset o, with x - cell in [0,1).  The polynomial piece is then
// known at compile time per tap, so each weight is a straight-line Horner chain
// with immediate coefficients -- no table lookup, no divergent index.  The Horner
// recurrence runs in the GLOBAL variable exactly like poly_util.hpp:19-27
// (highest power first), and the support tests of poly_util.hpp:32,42-47 are kept,
// so results differ from the reference only by FMA contraction (<= 1 ulp/step).
#pragma once
#include <cuda_runtime.h>

namespace spic {

#define SPIC_DI __device__ __forceinline__
#define SPIC_HDI __host__ __device__ __forceinline__

template <int N>
SPIC_HDI double horner(double x, const double (&c)[N]) {
  double s = c[0];
#pragma unroll
  for (int i = 1; i < N; ++i) s = fma(x, s, c[i]);
  return s;
}

// ---- 8th-order piecewise polynomial on [-2,2] ---------------------------------
struct InterpP8R2 {
  static constexpr int W = 2;        // interpolation_range (interpolation.cpp:20)
  static constexpr int NW1 = 4;      // W1 taps, offsets -1..2  (propagators.hpp:83-85)
  static constexpr int NWP = 3;      // Wp taps, offsets -1..1  (propagators.hpp:86-88)

  // piece p of W1 covers [p-2, p-1)  (interpolation.cpp:23-28)
  template <int P>
  static SPIC_HDI double w1_piece(double x) {
    if (P == 0) {
      const double c[9] = {15.0 / 1024, 15.0 / 128, 49.0 / 128, 21.0 / 32, 35.0 / 64, 0, 0, 1, 1};
      return horner(x, c);
    } else if (P == 1) {
      const double c[9] = {-15.0 / 1024, 15.0 / 128, 7.0 / 16, 21.0 / 32, 175.0 / 256,
                           0,            -105.0 / 128, 0,        337.0 / 512};
      return horner(x, c);
    } else if (P == 2) {
      const double c[9] = {-15.0 / 1024, -15.0 / 128, 7.0 / 16, -21.0 / 32, 175.0 / 256,
                           0,            -105.0 / 128, 0,        337.0 / 512};
      return horner(x, c);
    } else {
      const double c[9] = {15.0 / 1024, -15.0 / 128, 49.0 / 128, -21.0 / 32, 35.0 / 64, 0, 0, -1, 1};
      return horner(x, c);
    }
  }
  // piece p of Wp covers [p-1, p)  (interpolation.cpp:45-49)
  template <int P>
  static SPIC_HDI double wp_piece(double x) {
    if (P == 0) {
      const double c[8] = {15.0 / 128, 0, -21.0 / 128, 0, -35.0 / 128, 0, 105.0 / 128, 64.0 / 128};
      return horner(x, c);
    } else if (P == 1) {
      const double c[8] = {0,           105.0 / 128, -315.0 / 128, 420.0 / 128,
                           -315.0 / 128, 0,           105.0 / 128,  64.0 / 128};
      return horner(x, c);
    } else {
      const double c[8] = {-15.0 / 128, 105.0 / 128, -147.0 / 64, 105.0 / 32, -35.0 / 16, 0, 0, 1};
      return horner(x, c);
    }
  }
  // piece p of the running integral of Wp covers [p-1, p)  (interpolation.cpp:40-44)
  template <int P>
  static SPIC_HDI double iwp_piece(double x) {
    if (P == 0) {
      const double c[9] = {15.0 / 1024, 0, -28.0 / 1024, 0, -70.0 / 1024, 0, 420.0 / 1024,
                           512.0 / 1024, 175.0 / 1024};
      return horner(x, c);
    } else if (P == 1) {
      const double c[9] = {0,             120.0 / 1024, -420.0 / 1024, 672.0 / 1024, -630.0 / 1024,
                           0,             420.0 / 1024, 512.0 / 1024,  175.0 / 1024};
      return horner(x, c);
    } else {
      const double c[9] = {-15.0 / 1024, 120.0 / 1024, -392.0 / 1024, 672.0 / 1024, -560.0 / 1024,
                           0,            0,            1024.0 / 1024, 0};
      return horner(x, c);
    }
  }

  // tap T <-> offset o = T-1; argument a = x - (cell+o) lies in [1-T, 2-T]
  template <int T>
  static SPIC_HDI double w1_tap(double a) {
    const double v = w1_piece<3 - T>(a);
    return (a >= 2.0 || a <= -2.0) ? 0.0 : v;  // poly_util.hpp:32
  }
  template <int T>
  static SPIC_HDI double wp_tap(double a) {
    const double v = wp_piece<2 - T>(a);
    return (a >= 2.0 || a <= -1.0) ? 0.0 : v;
  }
  // running integral at a = s - (cell+o), s - cell in [0,1]  (poly_util.hpp:40-51)
  template <int T>
  static SPIC_HDI double iwp_tap(double a) {
    const double v = iwp_piece<2 - T>(a);
    return a >= 2.0 ? 1.0 : (a < -1.0 ? 0.0 : v);
  }

  // I_Wp(a, b) for the tap's node (propagators.hpp:178-186): difference of the running integral
  template <int T>
  static SPIC_HDI double iwp_seg(double a, double b) { return iwp_tap<T>(b) - iwp_tap<T>(a); }

  // In-cell tap forms (the binned kernels): the argument is KNOWN to lie in the closed range of
  // the tap's piece (particle inside its bin cell, segment inside its cell), so the support tests
  // of poly_util.hpp:32,42-47 can only fire at a piece end point -- and there every piece
  // polynomial evaluates EXACTLY (dyadic coefficients, dyadic argument) to the value the test
  // would return (0 or 1).  Same bits as w1_tap / wp_tap / iwp_tap, without DSETP/FSEL.
  template <int T>
  static SPIC_HDI double w1_in(double a) { return w1_piece<3 - T>(a); }
  template <int T>
  static SPIC_HDI double wp_in(double a) { return wp_piece<2 - T>(a); }
  template <int T>
  static SPIC_HDI double iwp_in(double a) { return iwp_piece<2 - T>(a); }
  template <int T>
  static SPIC_HDI double iwp_seg_in(double a, double b) { return iwp_in<T>(b) - iwp_in<T>(a); }

  // general-argument forms (diagnostics, tests): dynamic piece like the reference
  static SPIC_HDI double W1(double x) {
    if (x >= 2.0 || x <= -2.0) return 0.0;
    const int p = (int)floor(x) + 2;
    return p == 0 ? w1_piece<0>(x) : p == 1 ? w1_piece<1>(x) : p == 2 ? w1_piece<2>(x) : w1_piece<3>(x);
  }
  static SPIC_HDI double Wp(double x) {
    if (x >= 2.0 || x <= -1.0) return 0.0;
    const int p = (int)floor(x) + 1;
    return p == 0 ? wp_piece<0>(x) : p == 1 ? wp_piece<1>(x) : wp_piece<2>(x);
  }
  static SPIC_HDI double IWp_cdf(double q) {
    if (q >= 2.0) return 1.0;
    if (q < -1.0) return 0.0;
    const int p = (int)floor(q) + 1;
    return p == 0 ? iwp_piece<0>(q) : p == 1 ? iwp_piece<1>(q) : iwp_piece<2>(q);
  }
  static SPIC_HDI double I_Wp(double a, double b) { return IWp_cdf(b) - IWp_cdf(a); }
  static SPIC_HDI double IW1_cdf(double q) {  // interpolation.cpp:53-58, 73-75
    if (q >= 2.0) return 1.0;
    if (q < -2.0) return 0.0;
    const int p = (int)floor(q) + 2;
    if (p == 0) {
      const double c[10] = {5.0 / 3072, 15.0 / 1024, 7.0 / 128, 7.0 / 64, 7.0 / 64, 0, 0, 1.0 / 2, 1, 7.0 / 12};
      return horner(q, c);
    } else if (p == 1) {
      const double c[10] = {-5. / 3072, 15. / 1024, 1. / 16, 7. / 64, 35. / 256, 0, -35. / 128, 0, 337. / 512, 1.0 / 2};
      return horner(q, c);
    } else if (p == 2) {
      const double c[10] = {-5. / 3072, -15. / 1024, 1. / 16, -7. / 64, 35. / 256, 0, -35. / 128, 0, 337. / 512, 1.0 / 2};
      return horner(q, c);
    }
    const double c[10] = {5.0 / 3072, -15.0 / 1024, 7.0 / 128, -7.0 / 64, 7.0 / 64, 0, 0, -1.0 / 2, 1, 5.0 / 12};
    return horner(q, c);
  }
  static SPIC_HDI double I_W1(double a, double b) { return IW1_cdf(b) - IW1_cdf(a); }
};

// ---- piecewise linear on [-1,1] -------------------------------------------------
struct InterpPWL {
  static constexpr int W = 1;    // interpolation.cpp:89
  static constexpr int NW1 = 2;  // offsets 0..1
  static constexpr int NWP = 1;  // offset 0

  static SPIC_HDI double W1(double x) {  // interpolation.cpp:91-99
    const double fx = fabs(x);
    return fx >= 1.0 ? 0.0 : 1.0 - fx;
  }
  static SPIC_HDI double Wp(double x) { return (x < 1.0 && x >= 0.0) ? 1.0 : 0.0; }  // :134-141
  static SPIC_HDI double IWp_cdf(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }  // :143-153
  static SPIC_HDI double I_Wp(double a, double b) { return IWp_cdf(b) - IWp_cdf(a); }       // :155-157
  static SPIC_HDI double IW1_cdf(double x) {  // :118-128
    if (x > 1.0) return 1.0;
    if (x < -1.0) return 0.0;
    return -(x * fabs(x) - 2 * x - 1) * 0.5;
  }
  static SPIC_HDI double I_W1(double a, double b) { return IW1_cdf(b) - IW1_cdf(a); }  // :130-132

  template <int T>
  static SPIC_HDI double w1_tap(double a) { return W1(a); }
  template <int T>
  static SPIC_HDI double wp_tap(double a) { return Wp(a); }
  template <int T>
  static SPIC_HDI double iwp_tap(double a) { return IWp_cdf(a); }
  template <int T>
  static SPIC_HDI double iwp_seg(double a, double b) { return IWp_cdf(b) - IWp_cdf(a); }
  // In-cell tap forms (the binned kernels): the argument is KNOWN to lie in the tap's range -- tap 0 gets f in [0,1),
  // tap 1 gets f - 1 in [-1,0), the running integral gets s - cell in [0,1] -- so the range tests and the fabs of the
  // general forms are decided at compile time: 1 - |f| = 1 - f, 1 - |f - 1| = 1 + (f - 1), Wp = 1, clamp(a) = a.  Same
  // bits as the general forms for every such argument (tests/test_interp_host.py), without DSETP / FSEL: the PWL block
  // kernel is issue-bound (profiles/r02_ncu_axis_block_pwl_summary.txt).
  template <int T>
  static SPIC_HDI double w1_in(double a) { return T == 0 ? 1.0 - a : 1.0 + a; }
  template <int T>
  static SPIC_HDI double wp_in(double) { return 1.0; }
  template <int T>
  static SPIC_HDI double iwp_in(double a) { return a; }
  template <int T>
  static SPIC_HDI double iwp_seg_in(double a, double b) { return b - a; }
};

// Fill w1[0..NW1) and wp[0..NWP) for a particle at normalised coordinate x in cell c:
// w1[t] = W1(x - (c + t - W + 1)), wp[t] = Wp(x - (c + t - W + 1))   (propagators.hpp:138-165)
template <class I, int T = 0>
struct TapLoop {
  static SPIC_HDI void w1(double x, int c, double* out) {
    out[T] = I::template w1_tap<T>(x - (double)(c + T - I::W + 1));
    TapLoop<I, T + 1>::w1(x, c, out);
  }
};
template <class I>
struct TapLoop<I, 4> {
  static SPIC_HDI void w1(double, int, double*) {}
};

template <class I>
SPIC_HDI void eval_w1(double x, int c, double (&out)[I::NW1]) {
  if (I::NW1 == 4) {
    out[0] = I::template w1_tap<0>(x - (double)(c + 0 - I::W + 1));
    out[1] = I::template w1_tap<1>(x - (double)(c + 1 - I::W + 1));
    out[I::NW1 > 2 ? 2 : 0] = I::template w1_tap<2>(x - (double)(c + 2 - I::W + 1));
    out[I::NW1 > 3 ? 3 : 0] = I::template w1_tap<3>(x - (double)(c + 3 - I::W + 1));
  } else {
    out[0] = I::template w1_tap<0>(x - (double)(c + 0 - I::W + 1));
    out[1] = I::template w1_tap<1>(x - (double)(c + 1 - I::W + 1));
  }
}
template <class I>
SPIC_HDI void eval_wp(double x, int c, double (&out)[I::NWP]) {
  if (I::NWP == 3) {
    out[0] = I::template wp_tap<0>(x - (double)(c + 0 - I::W + 1));
    out[I::NWP > 1 ? 1 : 0] = I::template wp_tap<1>(x - (double)(c + 1 - I::W + 1));
    out[I::NWP > 2 ? 2 : 0] = I::template wp_tap<2>(x - (double)(c + 2 - I::W + 1));
  } else {
    out[0] = I::template wp_tap<0>(x - (double)(c + 0 - I::W + 1));
  }
}
// I[t] = I_Wp(s - cc, e - cc), cc = cell + t - W + 1   (propagators.hpp:178-186)
template <class I>
SPIC_HDI void eval_iwp(double s, double e, int cell, double (&out)[I::NWP]) {
  if (I::NWP == 3) {
    const double c0 = (double)(cell + 0 - I::W + 1), c1 = (double)(cell + 1 - I::W + 1),
                 c2 = (double)(cell + 2 - I::W + 1);
    out[0] = I::template iwp_seg<0>(s - c0, e - c0);
    out[I::NWP > 1 ? 1 : 0] = I::template iwp_seg<1>(s - c1, e - c1);
    out[I::NWP > 2 ? 2 : 0] = I::template iwp_seg<2>(s - c2, e - c2);
  } else {
    const double c0 = (double)(cell + 0 - I::W + 1);
    out[0] = I::template iwp_seg<0>(s - c0, e - c0);
  }
}

// ---- in-cell forms: f = x - cell in [0,1) is exact, so f - o has the bits of x - (cell + o) ----------
template <class I>
SPIC_HDI void eval_w1_in(double f, double (&out)[I::NW1]) {
  if (I::NW1 == 4) {
    out[0] = I::template w1_in<0>(f + 1.0);
    out[1] = I::template w1_in<1>(f);
    out[I::NW1 > 2 ? 2 : 0] = I::template w1_in<2>(f - 1.0);
    out[I::NW1 > 3 ? 3 : 0] = I::template w1_in<3>(f - 2.0);
  } else {
    out[0] = I::template w1_in<0>(f);
    out[1] = I::template w1_in<1>(f - 1.0);
  }
}
template <class I>
SPIC_HDI void eval_wp_in(double f, double (&out)[I::NWP]) {
  if (I::NWP == 3) {
    out[0] = I::template wp_in<0>(f + 1.0);
    out[I::NWP > 1 ? 1 : 0] = I::template wp_in<1>(f);
    out[I::NWP > 2 ? 2 : 0] = I::template wp_in<2>(f - 1.0);
  } else {
    out[0] = I::template wp_in<0>(f);
  }
}
// I[t] = I_Wp(s - cc, e - cc) for a segment [s,e] inside cell `cell`; hc = (double)cell
template <class I>
SPIC_HDI void eval_iwp_in(double s, double e, double hc, double (&out)[I::NWP]) {
  if (I::NWP == 3) {
    const double c0 = hc - 1.0, c2 = hc + 1.0;
    out[0] = I::template iwp_seg_in<0>(s - c0, e - c0);
    out[I::NWP > 1 ? 1 : 0] = I::template iwp_seg_in<1>(s - hc, e - hc);
    out[I::NWP > 2 ? 2 : 0] = I::template iwp_seg_in<2>(s - c2, e - c2);
  } else {
    out[0] = I::template iwp_seg_in<0>(s - hc, e - hc);
  }
}

}  // namespace spic

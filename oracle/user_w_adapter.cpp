// TEST INFRASTRUCTURE ONLY.  Links a user-W file (include/strugepic_user_w.h) into the REFERENCE the way the
// reference documents it: strong definitions of W1 / Wp / I_W1 / I_Wp / interpolation_range
// (include/strugepic_w.hpp:12-16 of MoPHA/strugepic) that take the place of the weak defaults of
// src/interpolation/interpolation.cpp.  Built by oracle/build_oracle.py: build_ref_user with -DWRANGE=<range>.
#include "strugepic_w.hpp"

#include "../include/strugepic_user_w.h"

#ifndef WRANGE
#error "define WRANGE = the user's interpolation range"
#endif

extern const int interpolation_range = WRANGE;
amrex::Real W1(amrex::Real x) { return spic_user_W1(x); }
amrex::Real Wp(amrex::Real x) { return spic_user_Wp(x); }
amrex::Real I_W1(amrex::Real a, amrex::Real b) { return spic_user_I_W1(a, b); }
amrex::Real I_Wp(amrex::Real a, amrex::Real b) { return spic_user_I_Wp(a, b); }

"""The user-W slot (include/strugepic_user_w.h, SPIC_INTERP_USER): the reference lets a user replace W1 / Wp /
I_W1 / I_Wp / interpolation_range by linking his own definitions over the weak defaults
(include/strugepic_w.hpp:12-16, src/interpolation/interpolation.cpp:10,14,20,89).  Here the same file is
(a) device-linked into libstrugepic_b200 and (b) linked into the REFERENCE (oracle/_ref/liboracle_ref_user.so),
and the two must agree -- on the host functions (CPU) and on every sub-flow and composition map (GPU).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle as ora
import strugepic_b200 as spic
import util

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

USER = spic.USER
TOL_STEP = 1e-11


def test_default_user_w_is_a_charge_conserving_pair():
    """csrc/user_w_default.cu (cubic B-spline): range, partition of unity, W1' = Wp(.+1) - Wp(.), integrals."""
    assert spic.interpolation_range(USER) == 2
    assert spic.W1(0.0, USER) == 2.0 / 3.0 and spic.W1(2.0, USER) == 0.0 and spic.Wp(0.5, USER) == 0.75
    assert spic.I_Wp(-1.0, 2.0, USER) == 1.0 and spic.I_W1(-2.0, 2.0, USER) == 1.0
    for x in np.linspace(0.0, 1.0, 21)[:-1]:
        assert abs(sum(spic.W1(x - i, USER) for i in range(-2, 4)) - 1) < 4e-16
        assert abs(sum(spic.Wp(x - i, USER) for i in range(-2, 4)) - 1) < 4e-16
    a, b = 0.21, 0.83  # the identity that makes the deposition charge conserving
    for i in range(-2, 3):
        lhs = spic.I_Wp(a - i, b - i, USER) - spic.I_Wp(a - i + 1, b - i + 1, USER)
        assert abs(lhs + (spic.W1(b - i, USER) - spic.W1(a - i, USER))) < 1e-15


def test_user_w_host_functions_match_the_reference_linked_with_the_same_file():
    g = np.load(os.path.join(HERE, "golden", "w_tables.npz"))
    xs = g["xs"]
    for name, fn, args in (("W1", spic.W1, lambda x: (x,)), ("Wp", spic.Wp, lambda x: (x,)),
                           ("I_Wp", spic.I_Wp, lambda x: (x, x + 0.37)), ("I_W1", spic.I_W1, lambda x: (x, x + 0.37))):
        got = np.array([fn(*args(x), interp=USER) for x in xs])
        assert np.max(np.abs(got - g["user_" + name])) <= 4e-16, name
    if ora.have_ref(USER):
        o = ora.RefOracle((4, 4, 4), interp=USER)
        assert o.W == 2 and o.W1(0.3) == pytest.approx(spic.W1(0.3, USER), abs=2e-16)


@pytest.mark.skipif(not ora.have_ref(USER), reason="needs /root/reference")
@pytest.mark.parametrize("name", sorted(make_golden.USER_CASES))
def test_user_golden_vectors_are_reproducible(name):
    c = make_golden.build_case(name)
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    o = ora.RefOracle(c["n_cell"], periodic=c["periodic"], interp=USER)
    util.load_state(o, c["E"], c["B"], c["parts"], c["q"], c["m"])
    util.run(o, c["schedule"])
    E, B, P = util.state_of(o)
    assert np.array_equal(E, g["E1"]) and np.array_equal(B, g["B1"]) and np.array_equal(P, g["P1"])


USER_PWL = r'''
// a user's restatement of the piecewise-linear pair: what `--user-w` takes
#include "strugepic_user_w.h"
#include <math.h>
SPIC_W_CONST int spic_user_interpolation_range = 1;
SPIC_W_FN double spic_user_W1(double x) { const double a = fabs(x); return a >= 1.0 ? 0.0 : 1.0 - a; }
SPIC_W_FN double spic_user_Wp(double x) { return (x >= 0.0 && x < 1.0) ? 1.0 : 0.0; }
static
#ifdef __CUDACC__
__host__ __device__
#endif
inline double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }
SPIC_W_FN double spic_user_I_Wp(double a, double b) { return clamp01(b) - clamp01(a); }
static
#ifdef __CUDACC__
__host__ __device__
#endif
inline double tri_cdf(double x) { return x > 1.0 ? 1.0 : (x < -1.0 ? 0.0 : -(x * fabs(x) - 2 * x - 1) * 0.5); }
SPIC_W_FN double spic_user_I_W1(double a, double b) { return tri_cdf(b) - tri_cdf(a); }
'''


@pytest.fixture(scope="module")
def user_pwl_library(tmp_path_factory):
    """`python -m strugepic_b200.build --user-w <file> --out <lib>`: a second library with the user's file in
    the slot (range 1 this time); the stock library is left alone."""
    from strugepic_b200 import build as spbuild
    d = tmp_path_factory.mktemp("userw")
    src = d / "my_pwl_w.cu"
    src.write_text(USER_PWL)
    out = os.path.join(spbuild.LIBDIR, "libstrugepic_b200_userpwl.so")
    lib = spbuild.build(user_w=str(src), out=out)
    assert lib == out and os.path.isfile(out)
    return out


def test_build_with_a_user_file_swaps_the_slot(user_pwl_library):
    lib = C.CDLL(user_pwl_library)
    for nm, na in (("spic_W1", 1), ("spic_Wp", 1), ("spic_I_Wp", 2), ("spic_I_W1", 2)):
        fn = getattr(lib, nm)
        fn.restype = C.c_double
        fn.argtypes = [C.c_int] + [C.c_double] * na
    assert lib.spic_interpolation_range(USER) == 1
    for x in np.linspace(-1.3, 1.3, 27):
        assert lib.spic_W1(USER, x) == spic.W1(x, spic.PWL) and lib.spic_Wp(USER, x) == spic.Wp(x, spic.PWL)
        assert lib.spic_I_Wp(USER, x, x + 0.4) == spic.I_Wp(x, x + 0.4, spic.PWL)
        assert lib.spic_I_W1(USER, x, x + 0.4) == spic.I_W1(x, x + 0.4, spic.PWL)
    assert lib.spic_W1(spic.P8R2, 0.0) == 0.658203125  # the shipped variants are still there
    assert spic.interpolation_range(USER) == 2         # and the stock library was not touched


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.USER_CASES))
def test_user_w_golden_vectors_gpu(name):
    c = make_golden.build_case(name)
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    s = spic.Simulation(c["n_cell"], periodic=c["periodic"], interp=USER)
    util.load_state(s, c["E"], c["B"], c["parts"], c["q"], c["m"])
    util.run(s, c["schedule"])
    nsteps = len(c["schedule"])
    util.compare_states((g["E1"], g["B1"], g["P1"]), util.state_of(s), TOL_STEP * nsteps, TOL_STEP * nsteps,
                        box=c["n_cell"])
    assert np.allclose(np.array(s.get_total_energy()), g["energy1"], rtol=1e-11, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
def test_user_w_every_subflow_against_the_reference(periodic):
    if not ora.have_ref(USER):
        pytest.skip("oracle/_ref/liboracle_ref_user.so missing")
    n_cell = (16, 9, 6)
    E, B = util.rng_fields(n_cell, 51)
    parts = util.plasma(n_cell, 6, 0.25, 51, periodic, 2)
    q, m = -1.0 / 6, 100.0 / 6
    o = ora.RefOracle(n_cell, periodic=periodic, interp=USER)
    s = spic.Simulation(n_cell, periodic=periodic, interp=USER)
    for t in (o, s):
        util.load_state(t, E, B, parts, q, m)
    ops = [("E", 0.3), ("axis", 0, 0.4), ("axis", 1, -0.4), ("axis", 2, 0.5), ("B", 0.7), ("axis", 2, -0.5),
           ("map", 1, 0.5), ("map", 2, 0.5), ("map", 4, 0.5)]
    for n, op in enumerate(ops, 1):
        util.apply(o, op)
        util.apply(s, op)
        util.compare_states(util.state_of(o), util.state_of(s), TOL_STEP * n * 3, TOL_STEP * n * 3, box=n_cell)
    assert np.allclose(s.number_density(), o.number_density(), rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_user_w_conserves_charge():
    """Discrete Gauss residual constant under all sub-flows with the user's pair (check #2 of north_star)."""
    n_cell = (12, 10, 8)
    s = spic.Simulation(n_cell, interp=USER)
    E, B = util.rng_fields(n_cell, 61, 0.3)
    parts = util.plasma(n_cell, 8, 0.2, 61)
    util.load_state(s, E, B, parts, -1.0 / 8, 100.0 / 8)
    g0 = s.gauss_residual()
    for _ in range(10):
        s.map(2, 0.5)
    g1 = s.gauss_residual()
    assert np.max(np.abs(g1 - g0)) < 1e-13 * max(1.0, np.max(np.abs(g0)))


@pytest.mark.gpu
@pytest.mark.parametrize("engine", [1, 0])
def test_user_pwl_library_equals_the_shipped_pwl(user_pwl_library, engine):
    """A library built with --user-w (PWL restated by a user, range 1) run in a subprocess against the stock
    library's PWL kernels: same particles and fields after maps of every order -- on the thread-per-particle engine
    and on the binned engine (fused axis block + particle-stream kernels instantiated over the user's functions)."""
    import subprocess
    code = r'''
import sys, numpy as np
sys.path[:0] = [%r, %r]
import strugepic_b200 as spic, util
n_cell = (9, 7, 6)
E, B = util.rng_fields(n_cell, 71, 0.4)
parts = util.plasma(n_cell, 5, 0.2, 71)
s = spic.Simulation(n_cell, interp=int(sys.argv[1]), engine=int(sys.argv[3]))
util.load_state(s, E, B, parts, -0.2, 20.0)
for order in (1, 2, 4):
    s.map(order, 0.5)
E1, B1, P1 = util.state_of(s)
P1 = P1[:, np.lexsort(np.round(P1[::-1], 9))]  # (the binned engine returns the particles in cell order)
np.savez(sys.argv[2], E=E1, B=B1, P=P1)
''' % (os.path.dirname(HERE), HERE)
    d = os.path.dirname(user_pwl_library)
    outs = []
    for tag, lib, interp in (("stock", None, spic.PWL), ("user", user_pwl_library, USER)):
        env = dict(os.environ)
        if lib:
            env["SPIC_B200_LIBRARY"] = lib
        out = os.path.join(d, "userpwl_%s.npz" % tag)
        r = subprocess.run([sys.executable, "-c", code, str(interp), out, str(engine)], env=env, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(np.load(out))
        os.remove(out)
    a, b = outs
    assert util.rel_err(b["E"], a["E"]) < 1e-12 and util.rel_err(b["B"], a["B"]) < 1e-12
    assert np.max(np.abs(b["P"] - a["P"])) < 1e-11

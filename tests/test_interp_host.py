"""CPU tests: the interpolation code the kernels inline (csrc/interp.cuh), evaluated on the host through
the C ABI, against the reference's W functions (golden table + oracle) and the in-cell tap forms against
the general forms BIT FOR BIT (they replace the support tests of poly_util.hpp:32,42-47 by exactness)."""
import os

import numpy as np
import pytest

import oracle as ora
import strugepic_b200 as spic

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("interp,tag", [(0, "p8"), (1, "pwl")])
def test_host_w_functions_match_reference_tables(interp, tag):
    """The kernels use FMA Horner chains, the reference (built without contraction) separate mul/add:
    the Horner recurrence runs in the global variable (|x| <= 2, partial sums of O(10)), so the two differ
    by up to 1e-14 absolute (cancellation near the support edge); exact at the dyadic points."""
    g = np.load(os.path.join(HERE, "golden", "w_tables.npz"))
    xs = g["xs"]
    tol = 2e-14 if interp == 0 else 0.0
    for name, fn, args in (("W1", spic.W1, lambda x: (x,)), ("Wp", spic.Wp, lambda x: (x,)),
                           ("I_Wp", spic.I_Wp, lambda x: (x, x + 0.37)), ("I_W1", spic.I_W1, lambda x: (x, x + 0.37))):
        got = np.array([fn(*args(x), interp=interp) for x in xs])
        assert np.max(np.abs(got - g[tag + "_" + name])) <= tol, name
    assert spic._lib.load().spic_interpolation_range(interp) == (2 if interp == 0 else 1)


def test_host_w_known_answers():
    assert spic.W1(0.0) == 0.658203125 and spic.W1(1.0) == 0.1708984375 and spic.W1(2.0) == 0.0
    assert spic.Wp(0.0) == 0.5 and spic.Wp(1.0) == 0.5 and spic.Wp(-1.0) == 0.0 and spic.Wp(2.0) == 0.0
    assert spic.I_Wp(-1.0, 2.0) == 1.0
    assert spic.W1(0.3, spic.PWL) == 0.7 and spic.I_Wp(0.2, 0.7, spic.PWL) == 0.7 - 0.2


@pytest.mark.parametrize("interp", [0, 1])
def test_in_cell_taps_are_bit_identical_to_the_general_forms(interp):
    lib = spic._lib.load()
    W = 2 if interp == 0 else 1
    nw1, nwp = 2 * W, 2 * W - 1
    rng = np.random.default_rng(5)
    fs = list(rng.random(400)) + [0.0, 0.5, 1.0 - 2.0 ** -53, 2.0 ** -60, 0.25, 0.75]
    o = ora.PortOracle((4, 4, 4), interp=interp)
    for cell in (0, 1, 7, 255):
        for f in fs:
            x = cell + f
            if not (np.floor(x) == cell):
                continue
            fx = x - cell
            for t in range(nw1):
                arg = x - float(cell + t - W + 1)
                assert lib.spic_tap_W1(interp, t, fx) == lib.spic_W1(interp, arg), (cell, f, t)
                assert abs(lib.spic_tap_W1(interp, t, fx) - o.W1(arg)) <= 2e-14
            for t in range(nwp):
                arg = x - float(cell + t - W + 1)
                assert lib.spic_tap_Wp(interp, t, fx) == lib.spic_Wp(interp, arg), (cell, f, t)
        # segments inside the cell, including both faces (closed interval)
        pts = [float(cell), float(cell + 1)] + [cell + f for f in fs[:60]]
        for s in pts[:12]:
            for e in pts:
                for t in range(nwp):
                    cc = float(cell + t - W + 1)
                    want = lib.spic_I_Wp(interp, s - cc, e - cc)
                    assert lib.spic_tap_IWp(interp, t, s, e, cell) == want, (cell, s, e, t)


def test_partition_of_unity_and_charge_identity_host():
    for x in np.linspace(0, 1, 33)[:-1]:
        assert abs(sum(spic.W1(x - i) for i in range(-2, 4)) - 1) < 4e-15
        assert abs(sum(spic.Wp(x - i) for i in range(-2, 4)) - 1) < 1e-14
    a, b = 0.21, 0.83
    for i in range(-2, 3):
        lhs = spic.I_Wp(a - i, b - i) - spic.I_Wp(a - i + 1, b - i + 1)
        assert abs(lhs + (spic.W1(b - i) - spic.W1(a - i))) < 3e-15

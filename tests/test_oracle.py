"""CPU tests (-m "not gpu"): the oracle itself.

 * the C port reproduces the golden vectors generated from the reference's own code
   (tests/golden/*.npz, made by tests/golden/make_golden.py);
 * where oracle/_ref exists (this container, or prebuilt on the GPU box) the port must
   agree with the unmodified reference BIT FOR BIT after every sub-flow;
 * the known answers / behaviours stated in SURVEY.md section 8(c) hold.
"""
import os
import sys

import numpy as np
import pytest

import oracle as ora
import util

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

have_ref = ora.have_ref(0) and ora.have_ref(1)
needs_ref = pytest.mark.skipif(not have_ref, reason="oracle/_ref not built (no /root/reference)")


# ---- W-function known answers (interpolation.cpp compiled alone) ------------
KAT_P8 = [
    ("W1", (0.0,), 0.658203125), ("W1", (1.0,), 0.1708984375), ("W1", (-1.0,), 0.1708984375),
    ("W1", (0.5,), 0.48120498657226562), ("W1", (-0.5,), 0.48120498657226562),
    ("W1", (1.5,), 0.018795013427734375), ("W1", (0.3,), 0.58860976938476561),
    ("W1", (-0.25,), 0.60906244814395905), ("W1", (2.0,), 0.0), ("W1", (-2.0,), 0.0),
    ("Wp", (0.0,), 0.5), ("Wp", (1.0,), 0.5), ("Wp", (0.5,), 0.7435302734375),
    ("Wp", (-0.5,), 0.12823486328125), ("Wp", (1.5,), 0.12823486328125), ("Wp", (0.3,), 0.7008444921875),
    ("Wp", (-1.0,), 0.0), ("Wp", (2.0,), 0.0),
    ("I_Wp", (-1.0, 2.0), 1.0), ("I_Wp", (0.2, 0.7), 0.35934402343749994),
    ("I_W1", (-2.0, 2.0), 1.0000000000000004), ("I_W1", (-0.3, 0.4), 0.43716891615071618),
]
KAT_PWL = [("W1", (0.3,), 0.7), ("Wp", (0.0,), 1.0), ("Wp", (0.999,), 1.0), ("Wp", (1.0,), 0.0),
           ("I_Wp", (0.2, 0.7), 0.5), ("I_W1", (-0.3, 0.4), 0.575)]


@pytest.mark.parametrize("interp,kats", [(0, KAT_P8), (1, KAT_PWL)])
def test_w_known_answers_port(interp, kats):
    o = ora.PortOracle((4, 4, 4), interp=interp)
    for name, args, want in kats:
        got = getattr(o, name)(*args)
        assert got == pytest.approx(want, abs=2e-16, rel=0), (name, args, got, want)


@needs_ref
@pytest.mark.parametrize("interp,kats", [(0, KAT_P8), (1, KAT_PWL)])
def test_w_known_answers_reference(interp, kats):
    o = ora.RefOracle((4, 4, 4), interp=interp)
    for name, args, want in kats:
        got = getattr(o, name)(*args)
        assert got == pytest.approx(want, abs=2e-16, rel=0), (name, args, got, want)


@pytest.mark.parametrize("interp,tag", [(0, "p8"), (1, "pwl")])
def test_w_tables_golden(interp, tag):
    g = np.load(os.path.join(HERE, "golden", "w_tables.npz"))
    o = ora.PortOracle((4, 4, 4), interp=interp)
    xs = g["xs"]
    assert np.array_equal(np.array([o.W1(x) for x in xs]), g[tag + "_W1"])
    assert np.array_equal(np.array([o.Wp(x) for x in xs]), g[tag + "_Wp"])
    assert np.array_equal(np.array([o.I_Wp(x, x + 0.37) for x in xs]), g[tag + "_I_Wp"])
    assert np.array_equal(np.array([o.I_W1(x, x + 0.37) for x in xs]), g[tag + "_I_W1"])


def test_w_structural_identities():
    o = ora.PortOracle((4, 4, 4), interp=0)
    for x in np.linspace(0.0, 1.0, 41)[:-1]:
        assert abs(sum(o.W1(x - i) for i in range(-2, 4)) - 1) < 4e-15
        assert abs(sum(o.Wp(x - i) for i in range(-2, 4)) - 1) < 1e-14
    # I_Wp(a-i,b-i) - I_Wp(a-i+1,b-i+1) = -(W1(b-i) - W1(a-i)): what makes deposition charge conserving
    a, b = 0.21, 0.83
    for i in range(-2, 3):
        lhs = o.I_Wp(a - i, b - i) - o.I_Wp(a - i + 1, b - i + 1)
        assert abs(lhs + (o.W1(b - i) - o.W1(a - i))) < 3e-15


def test_construct_segments():
    o = ora.PortOracle((4, 4, 4))
    assert o.construct_segments(5.2, 5.7)[0] == 1
    n, pts, idx = o.construct_segments(5.7, 6.1)
    assert (n, pts[1], idx) == (2, 6.0, (5, 6))
    n, pts, idx = o.construct_segments(5.1, 4.8)
    assert (n, pts[1], idx) == (2, 5.0, (5, 4))
    n, pts, idx = o.construct_segments(-0.1, 0.2)
    assert (n, pts[1], idx) == (2, 0.0, (-1, 0))


# ---- golden vectors (made from the reference) -----------------------------------------------
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_port_matches_golden(name):
    c = make_golden.build_case(name)
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    for k, v in (("E0", c["E"]), ("B0", c["B"]), ("P0", np.stack(c["parts"]))):
        assert np.array_equal(g[k], v), "golden inputs are not reproducible: " + k
    o = ora.PortOracle(c["n_cell"], periodic=c["periodic"], interp=c["interp"])
    util.load_state(o, c["E"], c["B"], c["parts"], c["q"], c["m"])
    util.run(o, c["schedule"])
    E, B, P = util.state_of(o)
    # same compiler flags (-O2 -ffp-contract=off) => bit identical
    assert np.array_equal(E, g["E1"]) and np.array_equal(B, g["B1"]) and np.array_equal(P, g["P1"])
    assert np.array_equal(np.array(o.energy()), g["energy1"])


@needs_ref
@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
def test_port_equals_reference_every_subflow(interp, periodic):
    n_cell = (12, 5, 4)
    W = 2 if interp == 0 else 1
    E, B = util.rng_fields(n_cell, 5)
    parts = util.plasma(n_cell, 3, 0.25, 5, periodic, W)
    q, m = -1.0 / 3, 100.0 / 3
    a = ora.PortOracle(n_cell, periodic=periodic, interp=interp)
    b = ora.RefOracle(n_cell, periodic=periodic, interp=interp)
    for o in (a, b):
        util.load_state(o, E, B, parts, q, m)
    ops = [("E", 0.3), ("axis", 0, 0.4), ("axis", 1, -0.4), ("axis", 2, 0.5), ("B", 0.7),
           ("source", 5, 2, 0.2, 0.3, 0.5, 1.5), ("map", 1, 0.5), ("map", 2, 0.5), ("map", 4, 0.5)]
    for op in ops:
        util.apply(a, op)
        util.apply(b, op)
        sa, sb = util.state_of(a), util.state_of(b)
        for x, y in zip(sa, sb):
            assert np.array_equal(x, y), op
        assert a.energy() == b.energy()


# ---- behaviours stated by the reference's decks (test/single_particle/*.input) ----------------
Q_E, M_E = -1.60217662e-19, 9.427127615688092e-16


def _single(o, pos, vel, E0, B0, n_cell):
    E = np.zeros((3, n_cell[2], n_cell[1], n_cell[0]))
    B = np.zeros_like(E)
    for c in range(3):
        E[c] = E0[c]
        B[c] = B0[c]
    util.load_state(o, E, B, [np.array([p]) for p in pos] + [np.array([v]) for v in vel], Q_E, M_E)


@pytest.mark.parametrize("kind", ["port"] + (["ref"] if have_ref else []))
def test_cyclotron_deck(kind):
    # cyclotron.input: one revolution in 1256.6 steps; SURVEY 8c known answers at steps 1, 2, 1256
    n = (12, 12, 12)
    o = util.make_oracle(kind, n, (1, 1, 1), 0)
    _single(o, (6, 4, 6), (0.01, 0.0, 0.01), (0, 0, 0), (0, 0, 58.8395), n)
    want = {1: (6.005, 4.0, 6.005, 0.01, 4.9999997388179764e-05, 0.01),
            2: (6.0099998750000125, 4.0000249999986943, 6.01, 0.0099997500000261162, 9.999874477655209e-05, 0.01)}
    for step in range(1, 1257):
        o.map(1, 0.5)
        if step in want:
            got = [float(t[0]) for t in o.get_particles()]
            assert got == pytest.approx(list(want[step]), rel=1e-14, abs=1e-18)
    x, y, z = [float(t[0]) for t in o.get_particles()[:3]]
    assert (x, y, z) == pytest.approx((5.9968209017141412, 4.000013001059183, 0.28000000000058278), rel=1e-12)
    kin = o.energy()[1]
    assert kin == pytest.approx(0.5 * M_E * 2e-4, rel=3e-5)


def test_reflection_deck():
    # reflection.input: 15x15x2, x not periodic, v_x flips at step 100 (x~12) and 280 (x~3)
    n = (15, 15, 2)
    o = ora.PortOracle(n, periodic=(0, 1, 1), interp=0)
    _single(o, (7, 7, 1), (0.1, 0.0, 0.0), (0, 0, 0), (0, 0, 0), n)
    flips, prev = [], 0.1
    for step in range(1, 300):
        o.map(1, 0.5)
        vx = float(o.get_particles()[3][0])
        if vx * prev < 0:
            flips.append((step, float(o.get_particles()[0][0])))
        prev = vx
    assert [f[0] for f in flips] == [100, 280]
    assert flips[0][1] == pytest.approx(12.0, abs=1e-12) and flips[1][1] == pytest.approx(3.0, abs=1e-12)


def test_break_deck_stops():
    # break.input: E_x decelerates v_x = 0.1 to rest in 3685.5 steps
    n = (12, 12, 12)
    o = ora.PortOracle(n, interp=0)
    _single(o, (6, 6, 6), (0.1, 0.0, 0.1), (0.31929995744680856, 0, 0), (0, 0, 0), n)
    for _ in range(3680):
        o.map(1, 0.5)
    vx = float(o.get_particles()[3][0])
    assert abs(vx) < 0.1 * 2e-3


def test_gauss_invariant_and_energy_envelope():
    # G(t) - G(0) stays at round-off under every sub-flow; H bounded (symplectic)
    n = (8, 8, 8)
    o = ora.PortOracle(n, interp=0)
    E = np.ones((3, 8, 8, 8))
    parts = util.plasma(n, 2, 0.01, 12345)
    util.load_state(o, E, E.copy(), parts, -1.0 / 2, 100.0 / 2)
    g0 = o.gauss()
    h0 = sum(o.energy())
    for _ in range(40):
        o.map(2, 0.5)
    drift = np.max(np.abs(o.gauss() - g0))
    assert drift < 2e-13 * max(1.0, np.max(np.abs(g0)))
    assert abs(sum(o.energy()) - h0) / h0 < 1e-3


def test_map4_is_three_map2_with_unit_coefficients():
    # hpp:578: 1/(2*l+1) is integer division => alpha = 1, beta = -1
    n = (6, 6, 6)
    E, B = util.rng_fields(n, 3, 0.3)
    parts = util.plasma(n, 2, 0.1, 3)
    a, b = ora.PortOracle(n), ora.PortOracle(n)
    for o in (a, b):
        util.load_state(o, E, B, parts, -0.5, 50.0)
    a.map(4, 0.5)
    for dt in (0.5, -0.5, 0.5):
        b.map(2, dt)
    for x, y in zip(util.state_of(a), util.state_of(b)):
        assert np.array_equal(x, y)
    c = ora.PortOracle(n)
    util.load_state(c, E, B, parts, -0.5, 50.0)
    c.map(4, 0.3, yoshida=True)
    assert not np.array_equal(util.state_of(c)[0], util.state_of(a)[0])


@needs_ref
@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
def test_number_density_port_matches_reference(interp, periodic):
    """get_particle_number_density<W> (include/strugepic_util.hpp:30-85): port == reference bit for bit,
    and the deposit conserves the particle count (sum_i Wp(x - i) = 1) on a periodic box."""
    n_cell = (9, 7, 5)
    W = 2 if interp == 0 else 1
    parts = util.plasma(n_cell, 5, 0.2, 61, periodic, W)
    dens = []
    for kind in ("ref", "port"):
        o = util.make_oracle(kind, n_cell, periodic, interp)
        z = np.zeros((3, 5, 7, 9))
        util.load_state(o, z, z, parts, -1.0, 1.0)
        dens.append(o.number_density())
    assert np.array_equal(dens[0], dens[1])
    if all(periodic):
        assert abs(dens[0].sum() - len(parts[0])) < 1e-9 * len(parts[0])

"""GPU parity tests (-m gpu): the CUDA library, called through the C ABI, against the oracle.

Tolerances.  The kernels use FMA contraction and factorised stencil sums, and deposition is
a floating-point reduction whose order differs from the reference's serial particle loop, so
results are not bit-identical.  north_star budgets 1e-11 relative per step; measured
differences are ~1e-15 per sub-flow.  We assert TOL_STEP = 1e-11 per map step (fields
relative to max|F|, positions absolute in cells, velocities relative to max|v|).
"""
import os
import sys

import numpy as np
import pytest

import oracle as ora
import util

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

TOL_STEP = 1e-11
ENGINES = [0, 1]  # SPIC_ENGINE_BINNED, SPIC_ENGINE_DIRECT


def spic():
    import strugepic_b200
    return strugepic_b200


def gpu_sim(c, engine, **kw):
    s = spic().Simulation(c["n_cell"], periodic=c["periodic"], interp=c["interp"], engine=engine, **kw)
    util.load_state(s, c["E"], c["B"], c["parts"], c["q"], c["m"])
    return s


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_golden_vectors(name, engine):
    """Committed vectors generated from the reference's own code (tests/golden/make_golden.py)."""
    c = make_golden.build_case(name)
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    s = gpu_sim(c, engine)
    util.run(s, c["schedule"])
    nsteps = len(c["schedule"])
    errs = util.compare_states((g["E1"], g["B1"], g["P1"]), util.state_of(s), TOL_STEP * nsteps, TOL_STEP * nsteps,
                               box=c["n_cell"])
    en = np.array(s.get_total_energy())
    assert np.allclose(en, g["energy1"], rtol=1e-11, atol=0), (en, g["energy1"])
    print(name, errs)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
def test_every_subflow_against_oracle(interp, periodic, engine):
    n_cell = (16, 9, 6)
    W = 2 if interp == 0 else 1
    E, B = util.rng_fields(n_cell, 21)
    parts = util.plasma(n_cell, 6, 0.25, 21, periodic, W)
    q, m = -1.0 / 6, 100.0 / 6
    o = ora.best_oracle(n_cell, periodic=periodic, interp=interp)
    s = spic().Simulation(n_cell, periodic=periodic, interp=interp, engine=engine)
    for t in (o, s):
        util.load_state(t, E, B, parts, q, m)
    ops = [("E", 0.3), ("axis", 0, 0.4), ("axis", 1, -0.4), ("axis", 2, 0.5), ("B", 0.7),
           ("source", 5, 2, 0.2, 0.3, 0.5, 1.5), ("axis", 2, -0.5), ("axis", 0, 0.25), ("E", -0.3)]
    for n, op in enumerate(ops, 1):
        util.apply(o, op)
        util.apply(s, op)
        util.compare_states(util.state_of(o), util.state_of(s), TOL_STEP * n, TOL_STEP * n, box=n_cell)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("order", [1, 2, 4])
def test_maps_many_cell_crossings(order, engine):
    """32^3 x 8 ppc, hot plasma (10 % of the particles change cell per sub-flow), 5 steps."""
    n_cell = (32, 32, 32)
    E, B = util.rng_fields(n_cell, 31, 0.2)
    parts = util.plasma(n_cell, 8, 0.2, 31)
    q, m = -1.0 / 8, 100.0 / 8
    o = ora.best_oracle(n_cell, interp=0)
    s = spic().Simulation(n_cell, interp=0, engine=engine)
    for t in (o, s):
        util.load_state(t, E, B, parts, q, m)
    for _ in range(5):
        o.map(order, 0.5)
        s.map(order, 0.5)
    util.compare_states(util.state_of(o), util.state_of(s), TOL_STEP * 5, TOL_STEP * 5, box=n_cell)
    assert s.num_particles() == len(parts[0])


@pytest.mark.parametrize("engine", ENGINES)
def test_tiny_and_ragged_boxes(engine):
    """Edge shapes of the reference's own decks: 4x4x1 (energy.input), odd sizes, guard > box."""
    for n_cell, ng in (((4, 4, 1), 3), ((5, 3, 2), 2), ((15, 15, 2), 3)):
        E, B = util.rng_fields(n_cell, 41)
        parts = util.plasma(n_cell, 5, 0.2, 41)
        o = ora.best_oracle(n_cell, interp=0, ng=ng)
        s = spic().Simulation(n_cell, interp=0, ng=ng, engine=engine)
        for t in (o, s):
            util.load_state(t, E, B, parts, -0.2, 20.0)
        for _ in range(3):
            o.map(2, 0.5)
            s.map(2, 0.5)
        util.compare_states(util.state_of(o), util.state_of(s), 3 * TOL_STEP, 3 * TOL_STEP, box=n_cell)


@pytest.mark.parametrize("engine", ENGINES)
def test_empty_species_and_field_only(engine):
    """No particles: field_only schedule with source + MABC (examples/field_only/main.cpp:142-145)."""
    n_cell = (64, 4, 4)
    o = ora.best_oracle(n_cell, periodic=(0, 1, 1), interp=0, ng=3)
    s = spic().Simulation(n_cell, periodic=(0, 1, 1), interp=0, ng=3, engine=engine)
    z = np.zeros((3, 4, 4, 64))
    for t in (o, s):
        util.load_state(t, z, z, [np.zeros(0)] * 6, -1.0, 1.0)
    for step in range(120):
        for op in (("E", 0.25), ("source", 4, 1, 0.1, 0.3, 0.5, 0.5 * step), ("B", 0.5), ("E", 0.25)):
            util.apply(o, op)
        s.field_only_step(4, 1, 0.1, 0.3, 0.5, step)
    Eo, Bo, _ = util.state_of(o)
    Es, Bs, _ = util.state_of(s)
    assert util.rel_err(Es, Eo) < 1e-12 and util.rel_err(Bs, Bo) < 1e-12
    assert np.max(np.abs(Eo)) > 1e-2
    assert s.get_total_energy()[0] == pytest.approx(o.energy()[0], rel=1e-12)


@pytest.mark.parametrize("periodic", [(0, 1, 1), (1, 1, 1), (1, 0, 1)])
def test_field_only_at_scale(periodic):
    """BASELINE configs[2] shape at a size the oracle runs in seconds: 256 x 32 x 32 vacuum Maxwell, soft plane-wave
    source, MABC on the x faces (examples/field_only/main.cpp:142-145).  One launch per sub-flow here (source and MABC
    folded into the curl sweeps, periodic directions wrapped inside them, no guard refresh): equal to the oracle's
    separate passes to round-off; also fully periodic, and with a y wall (MABC_bad<X> still blends the x faces of a
    periodic x then: hpp:516-523 applies it whenever the box is not ALL periodic)."""
    n_cell = (256, 32, 32)
    o = ora.best_oracle(n_cell, periodic=periodic, interp=0, ng=3)
    s = spic().Simulation(n_cell, periodic=periodic, interp=0, ng=3)
    E, B = util.rng_fields(n_cell, 97, 1e-3)
    for t in (o, s):
        util.load_state(t, E, B, [np.zeros(0)] * 6, -1.0, 1.0)
    pos, comp, E0, omega, dt = 32, 1, 0.1, 0.3, 0.5
    for step in range(60):
        for op in (("E", dt / 2), ("source", pos, comp, E0, omega, dt, dt * step), ("B", dt), ("E", dt / 2)):
            util.apply(o, op)
        s.field_only_step(pos, comp, E0, omega, dt, step)
    Eo, Bo, _ = util.state_of(o)
    Es, Bs, _ = util.state_of(s)
    assert util.rel_err(Es, Eo) < 1e-12 and util.rel_err(Bs, Bo) < 1e-12
    assert np.max(np.abs(Eo)) > 1e-2
    # the same sub-flows called one by one (spic_source as its own launch) agree with the folded step
    s2 = spic().Simulation(n_cell, periodic=periodic, interp=0, ng=3)
    util.load_state(s2, E, B, [np.zeros(0)] * 6, -1.0, 1.0)
    for step in range(60):
        for op in (("E", dt / 2), ("source", pos, comp, E0, omega, dt, dt * step), ("B", dt), ("E", dt / 2)):
            util.apply(s2, op)
    E2, B2, _ = util.state_of(s2)
    assert util.rel_err(E2, Es) < 1e-13 and util.rel_err(B2, Bs) < 1e-13


@pytest.mark.parametrize("n_cell", [(64, 16, 8), (37, 9, 5), (256, 32, 32)])
def test_curl_sweeps_with_tma_tiles(n_cell):
    """Option curl_tma: the curl sweeps of a periodic box stage their S tiles with ONE 4-D tensor-map box per block
    (UTMALDG + mbarrier) and read the periodic neighbours from the guards.  Same arithmetic as the plain sweep:
    bit-identical fields, on boxes that are and are not multiples of the 32 x 8 x 4 tile, incl. the double Theta_E sweep
    of the field-only step and a PIC step."""
    E, B = util.rng_fields(n_cell, 99, 0.5)
    out = []
    for tma in (0, 1):
        s = spic().Simulation(n_cell, interp=0)
        s.set_option("curl_tma", tma)
        s.set_field(0, E)   # (spic_set_field refreshes the guards: the first sweeps below find them valid)
        s.set_field(1, B)
        s.set_option("time_kernels", 1)
        for k in range(4):
            s.G_Theta_E(0.3)
            s.set_field(1, s.get_field(1))  # guards of B valid again: Theta_B takes the tiled sweep when tma = 1
            s.G_Theta_B(0.4)
            s.set_field(0, s.get_field(0))
        for k in range(5):  # periodic box, no species: pending half + leading half in one sweep (dt, dt2)
            s.field_only_step(3, 1, 0.1, 0.3, 0.5, k)
        out.append((s.get_field(0), s.get_field(1)))
        s.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    if n_cell[0] == 37:  # and inside a PIC step, against the oracle
        parts = util.plasma(n_cell, 6, 0.1, 99)
        o = ora.best_oracle(n_cell, interp=0)
        s = spic().Simulation(n_cell, interp=0)
        s.set_option("curl_tma", 1)
        for t in (o, s):
            util.load_state(t, E, B, parts, -1.0 / 6, 100.0 / 6)
            t.map(4, 0.5)
            t.map(2, 0.5)
        util.compare_states(util.state_of(o), util.state_of(s), 4 * TOL_STEP, 4 * TOL_STEP, box=n_cell)


@pytest.mark.parametrize("engine", ENGINES)
def test_single_particle_decks(engine):
    """cyclotron.input known answers and reflection.input flip steps."""
    Q_E, M_E = -1.60217662e-19, 9.427127615688092e-16
    s = spic().Simulation((12, 12, 12), interp=0, ng=3, engine=engine)
    B = np.zeros((3, 12, 12, 12))
    B[2] = 58.8395
    s.set_field(0, np.zeros_like(B))
    s.set_field(1, B)
    s.add_species(Q_E, M_E, [6.0], [4.0], [6.0], [0.01], [0.0], [0.01])
    s.Theta_map1(0.5)
    got = [float(t[0]) for t in s.get_particles()]
    assert got == pytest.approx([6.005, 4.0, 6.005, 0.01, 4.9999997388179764e-05, 0.01], rel=1e-13, abs=1e-18)
    for _ in range(1255):
        s.Theta_map1(0.5)
    x, y, z = [float(t[0]) for t in s.get_particles()[:3]]
    assert (x, y, z) == pytest.approx((5.9968209017141412, 4.000013001059183, 0.28000000000058278), rel=1e-10)

    r = spic().Simulation((15, 15, 2), periodic=(0, 1, 1), interp=0, ng=3, engine=engine)
    zf = np.zeros((3, 2, 15, 15))
    r.set_field(0, zf)
    r.set_field(1, zf)
    r.add_species(Q_E, M_E, [7.0], [7.0], [1.0], [0.1], [0.0], [0.0])
    flips, prev = [], 0.1
    for step in range(1, 300):
        r.Theta_map1(0.5)
        vx = float(r.get_particles()[3][0])
        if vx * prev < 0:
            flips.append(step)
        prev = vx
    assert flips == [100, 280]


@pytest.mark.parametrize("engine", ENGINES)
def test_gauss_law_and_energy_at_scale(engine):
    """Check #2 and #3 of north_star on the energy_conservation config (64^3, 8 ppc, W8, map2):
    the discrete Gauss residual must not move beyond round-off and H must stay in the reference's
    envelope (map2 within [-3.12e-4, 0] relative)."""
    n_cell = (64, 64, 64)
    s = spic().Simulation(n_cell, interp=0, engine=engine)
    s.set_uniform_field(0, [1, 1, 1])
    s.set_uniform_field(1, [1, 1, 1])
    s.add_particle_density_uniform(8, 100.0, -1.0, 0.01, seed=12345)
    assert s.num_particles() == 64 ** 3 * 8
    g0 = s.gauss_residual()
    h0 = sum(s.get_total_energy())
    for _ in range(20):
        s.Theta_map2(0.5)
    g1 = s.gauss_residual()
    h1 = sum(s.get_total_energy())
    drift = float(np.max(np.abs(g1 - g0)))
    assert drift < 1e-12 * max(1.0, float(np.max(np.abs(g0)))), drift
    assert -4e-4 < (h1 - h0) / h0 < 1e-6, (h1 - h0) / h0
    s.Theta_map4(0.5)
    assert float(np.max(np.abs(s.gauss_residual() - g0))) < 1e-12 * max(1.0, float(np.max(np.abs(g0))))


@pytest.mark.parametrize("engine", ENGINES)
def test_device_loader_is_the_numpy_twin(engine):
    from strugepic_b200 import synthetic
    n_cell = (10, 7, 5)
    s = spic().Simulation(n_cell, interp=0, engine=engine)
    s.add_particle_density_uniform(4, 100.0, -1.0, 0.01, seed=777)
    got = np.stack(s.get_particles())
    want = np.stack(synthetic.uniform_plasma(n_cell, 4, 0.01, 777))
    got = got[:, util.match_particles(want, got)]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("engine", ENGINES)
def test_density_profile_loader_is_the_numpy_twin(engine):
    """spic_load_density_plasma (add_particle_density with a profile, util.cpp:267-311): counts per cell =
    int(profile * ppc_max), including profiles above 1 (simple_line_density) and empty cells; bit-identical
    to synthetic.density_plasma, and equal to the uniform loader when the profile is 1."""
    from strugepic_b200 import synthetic
    n_cell = (48, 3, 4)
    for prof, ppc in ((synthetic.simple_line_density, 5), (lambda n, i, j, k: np.exp(-((i - 24.0) / 8) ** 2) + 0.0 * (j + k), 7),
                      (lambda n, i, j, k: 0.5 + 0.25 * ((i + j + k) % 3), 6)):
        s = spic().Simulation(n_cell, interp=0, engine=engine)
        s.add_particle_density(prof, ppc, 100.0, -1.0, 0.02, seed=99)
        want = np.stack(synthetic.density_plasma(n_cell, prof, ppc, 0.02, 99))
        got = np.stack(s.get_particles())
        assert got.shape == want.shape and want.shape[1] > 0
        got = got[:, util.match_particles(want, got)]
        assert np.array_equal(got, want)
        counts, _ = synthetic.density_counts(n_cell, prof, ppc)
        cell = (np.floor(got[2]).astype(int) * n_cell[1] + np.floor(got[1]).astype(int)) * n_cell[0] + \
            np.floor(got[0]).astype(int)
        assert np.array_equal(np.bincount(cell, minlength=counts.size), counts.ravel())
        s.close()
    s = spic().Simulation(n_cell, interp=0, engine=engine)
    s.add_particle_density(synthetic.uniform_density, 4, 100.0, -1.0, 0.01, seed=777)
    want = np.stack(synthetic.uniform_plasma(n_cell, 4, 0.01, 777))
    got = np.stack(s.get_particles())
    assert np.array_equal(got[:, util.match_particles(want, got)], want)


@pytest.mark.parametrize("engine", ENGINES)
def test_energy_tracks_oracle(engine):
    n_cell = (16, 16, 16)
    parts = util.plasma(n_cell, 8, 0.01, 12345)
    E = np.ones((3, 16, 16, 16))
    o = ora.best_oracle(n_cell, interp=0)
    s = spic().Simulation(n_cell, interp=0, engine=engine)
    for t in (o, s):
        util.load_state(t, E, E.copy(), parts, -1.0 / 8, 100.0 / 8)
    for step in range(30):
        o.map(2, 0.5)
        s.map(2, 0.5)
        ho, hs = sum(o.energy()), sum(s.get_total_energy())
        assert abs(hs - ho) / ho < 1e-9, step


@pytest.mark.parametrize("engine", ENGINES)
def test_checkpoint_roundtrip(tmp_path, engine):
    n_cell = (12, 8, 6)
    c = dict(n_cell=n_cell, periodic=(1, 1, 1), interp=0, q=-0.25, m=25.0)
    c["E"], c["B"] = util.rng_fields(n_cell, 51)
    c["parts"] = util.plasma(n_cell, 4, 0.2, 51)
    a = gpu_sim(c, engine)
    a.Theta_map2(0.5)
    a.checkpoint(tmp_path / "ck.bin")
    b = spic().Simulation(n_cell, interp=0, engine=engine)
    b.restart(tmp_path / "ck.bin")
    for t in (a, b):
        t.Theta_map2(0.5)
    util.compare_states(util.state_of(a), util.state_of(b), 1e-13, 1e-13, box=n_cell)
    with pytest.raises(spic().SpicError):
        spic().Simulation((12, 8, 7), interp=0).restart(tmp_path / "ck.bin")


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
def test_number_density_and_plot_file(tmp_path, interp, periodic, engine):
    """get_particle_number_density<W> + SimulationIO::write<W>(step) (include/strugepic_util.hpp:30-85, 133-143)."""
    n_cell = (9, 7, 5)
    W = 2 if interp == 0 else 1
    E, B = util.rng_fields(n_cell, 61)
    parts = util.plasma(n_cell, 5, 0.2, 61, periodic, W)
    o = ora.best_oracle(n_cell, periodic=periodic, interp=interp)
    s = spic().Simulation(n_cell, periodic=periodic, interp=interp, engine=engine)
    for t in (o, s):
        util.load_state(t, E, B, parts, -0.2, 20.0)
    want = o.number_density()
    got = s.number_density()
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
    s.write_plot(tmp_path / "plt0.spic")
    plt = spic().read_plot(tmp_path / "plt0.spic")
    assert plt["n_cell"] == n_cell and plt["n"] == n_cell and plt["interp"] == interp
    assert np.array_equal(plt["E"], E) and np.array_equal(plt["B"], B)
    assert np.allclose(plt["n_density"], got, rtol=1e-13, atol=0)  # atomics: the order of additions varies


@pytest.mark.parametrize("variant", [2, 3])
def test_kernel_variants_agree_with_oracle(variant):
    """Both generations of the single-sub-flow kernels stay covered (option axis_kernel / pushve_kernel: 2 = warp per
    cell, 3 = particle stream), with the low-count kernels off so that the named ones run."""
    n_cell = (16, 12, 8)
    E, B = util.rng_fields(n_cell, 71, 0.3)
    parts = util.plasma(n_cell, 40, 0.15, 71)   # > 32 per cell: several batches per bin
    for interp in (0, 1):
        o = ora.best_oracle(n_cell, interp=interp)
        s = spic().Simulation(n_cell, interp=interp)
        s.set_option("axis_kernel", variant)
        s.set_option("pushve_kernel", variant)
        s.set_option("pair_kernel", 0)
        for t in (o, s):
            util.load_state(t, E, B, parts, -1.0 / 40, 100.0 / 40)
        for t in (o, s):
            t.map(2, 0.5)
            t.map(1, 0.5)
        util.compare_states(util.state_of(o), util.state_of(s), 2 * TOL_STEP, 2 * TOL_STEP, box=n_cell)


@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_fused_axis_block_and_reference_schedule(fuse, interp, mode):
    """Theta_map2/4 with the six position sub-flows of every map2 fused into one pass (option fuse = 1, the
    default on periodic boxes) and with the reference's launch-per-sub-flow schedule (fuse = 0) against the
    oracle; 40 ppc => cells with two batches, v_th = 0.08 => ~10 % of the particles leave their cell inside a
    block and are finished by the general per-particle code (k_axis_continue); both Theta_map4 coefficient modes."""
    n_cell = (12, 10, 7)
    E, B = util.rng_fields(n_cell, 77, 0.3)
    parts = util.plasma(n_cell, 40, 0.08, 77)
    q, m = -1.0 / 40, 100.0 / 40
    dt = 0.5 if mode == 0 else 0.3  # Yoshida: |beta| dt must stay below the vacuum CFL limit
    # (the reference's own Theta_map4 only knows alpha = 1, beta = -1: Yoshida is checked against the C port)
    o = ora.best_oracle(n_cell, interp=interp) if mode == 0 else util.make_oracle("port", n_cell, (1, 1, 1), interp)
    s = spic().Simulation(n_cell, interp=interp, map4_mode=mode)
    s.set_option("fuse", fuse)
    s.set_option("time_kernels", 1)
    for t in (o, s):
        util.load_state(t, E, B, parts, q, m)
    for _ in range(2):
        o.map(4, dt, yoshida=mode == 1)
        s.map(4, dt)
    for _ in range(2):
        o.map(2, dt)
        s.map(2, dt)
    errs = util.compare_states(util.state_of(o), util.state_of(s), TOL_STEP * 8, TOL_STEP * 8, box=n_cell)
    kt = s.kernel_times()
    assert kt["axis_block"][1] == (8 if fuse else 0), kt
    assert kt["theta_axis"][1] == (0 if fuse else 48), kt
    assert s.num_particles() == len(parts[0])
    print(fuse, interp, mode, errs, kt)


@pytest.mark.parametrize("tma", [0, 1])
@pytest.mark.parametrize("ppc", [1, 7, 8, 33, 64, 65])
def test_fused_block_batch_shapes(ppc, tma):
    """The fused axis block over bins of 1 ... 65+ particles: empty lanes, exactly full batches, a cell of 65
    (32 + 32 + 1), both W, warm plasma so that the counts drift apart and particles are ejected to the continuation;
    Gauss residual constant to round-off.  tma = 1: batches staged with cp.async.bulk + a 4-D tensor map (odd row
    lengths are rounded up to 16 bytes) instead of cp.async -- same results to the bit."""
    for interp, n_cell, vth, order in ((0, (8, 6, 5), 0.1, 4), (1, (9, 7, 3), 0.15, 2), (0, (16, 2, 2), 0.05, 2)):
        E, B = util.rng_fields(n_cell, 5, 0.3)
        parts = util.plasma(n_cell, ppc, vth, 5)
        o = ora.best_oracle(n_cell, interp=interp)
        s = spic().Simulation(n_cell, interp=interp)
        s.set_option("time_kernels", 1)
        s.set_option("tma", tma)
        for t in (o, s):
            util.load_state(t, E, B, parts, -1.0 / ppc, 100.0 / ppc)
        g0 = s.gauss_residual()
        for _ in range(3):
            o.map(order, 0.5)
            s.map(order, 0.5)
        errs = util.compare_states(util.state_of(o), util.state_of(s), 3 * TOL_STEP, 3 * TOL_STEP, box=n_cell)
        assert s.num_particles() == len(parts[0])
        assert s.kernel_times()["axis_block"][1] == (9 if order == 4 else 3)
        drift = float(np.max(np.abs(s.gauss_residual() - g0)))
        assert drift < 1e-12 * max(1.0, float(np.max(np.abs(g0)))), drift
        print(interp, n_cell, ppc, errs, drift)
        s.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_two_species(engine):
    """Electrons + ions (two (q, m) pairs: the reference carries q, m per particle, defs.hpp:22-49): the species
    share the mover list, the continuation buffers and the field deposits.  State vs the oracle, and the Gauss
    residual (rho of BOTH species) against the port's."""
    n_cell = (10, 8, 6)
    E, B = util.rng_fields(n_cell, 83, 0.3)
    el = util.plasma(n_cell, 36, 0.1, 83)
    io = util.plasma(n_cell, 20, 0.02, 84)
    qe, me, qi, mi = -1.0 / 36, 100.0 / 36, 1.0 / 20, 1836.0 / 20
    allp = [np.concatenate([a, b]) for a, b in zip(el, io)]
    qa = np.concatenate([np.full(len(el[0]), qe), np.full(len(io[0]), qi)])
    ma = np.concatenate([np.full(len(el[0]), me), np.full(len(io[0]), mi)])
    o = util.make_oracle("port", n_cell, (1, 1, 1), 0)
    o.set_field(0, E)
    o.set_field(1, B)
    o.set_particles(*allp, qa, ma)
    s = spic().Simulation(n_cell, interp=0, engine=engine)
    s.set_field(0, E)
    s.set_field(1, B)
    s.add_species(qe, me, *el)
    s.add_species(qi, mi, *io)
    go, gs = o.gauss(), s.gauss_residual()
    assert np.max(np.abs(gs - go)) < 1e-12 * np.max(np.abs(go))
    for order in (2, 4, 1):
        o.map(order, 0.5)
        s.map(order, 0.5)
    Eo, Bo = o.get_field(0), o.get_field(1)
    Po = np.stack(o.get_particles())
    Ps = np.concatenate([np.stack(s.get_particles(0)), np.stack(s.get_particles(1))], axis=1)
    errs = util.compare_states((Eo, Bo, Po), (s.get_field(0), s.get_field(1), Ps), 3 * TOL_STEP, 3 * TOL_STEP, box=n_cell)
    assert s.num_particles(0) == len(el[0]) and s.num_particles(1) == len(io[0])
    go1, gs1 = o.gauss(), s.gauss_residual()
    assert np.max(np.abs(gs1 - go1)) < 1e-11 * np.max(np.abs(go1))
    assert np.max(np.abs(gs1 - gs)) < 1e-12 * max(1.0, float(np.max(np.abs(gs))))
    ho, hs = sum(o.energy()), sum(s.get_total_energy())
    assert abs(hs - ho) <= 1e-11 * abs(ho)
    print(engine, errs)


def _sorted_cols(P):
    P = np.stack(P)
    return P[:, np.lexsort(P[::-1])]


@pytest.mark.parametrize("engine", ENGINES)
def test_reupload_reuses_the_particle_store(engine):
    """spic_set_particles on a species that already holds particles (what a caller does that keeps its particles
    on the host and hands them over every step): the bin arrays, the upload list and the permutation of the previous
    call are reused when they fit (csrc/particles_binned.cu: engine_upload).  Every upload must replace the whole
    store -- smaller, larger, the same size, after one map and after two (the retained list is released by the second)
    -- and a step after the n-th upload must equal the step of a context that was loaded once."""
    n_cell = (12, 10, 8)
    E, B = util.rng_fields(n_cell, 91, 0.3)
    sets = {k: util.plasma(n_cell, ppc, 0.1, 91 + k) for k, ppc in enumerate((20, 31, 7, 20))}
    q, m = -1.0 / 20, 1.0 / 20

    def fresh(parts):
        t = spic().Simulation(n_cell, interp=0, engine=engine)
        t.set_field(0, E)
        t.set_field(1, B)
        t.add_species(q, m, *parts)
        return t

    s = fresh(sets[0])
    for k in (1, 2, 3, 0):  # larger (new arrays), much smaller (arrays far too large), the first size again
        s.set_particles(0, *sets[k])
        assert s.num_particles(0) == len(sets[k][0])
        assert np.array_equal(_sorted_cols(s.get_particles(0)), _sorted_cols(sets[k]))
    ref = fresh(sets[0])
    for t in (s, ref):
        t.set_field(0, E)
        t.set_field(1, B)
        t.map(2, 0.5)
    util.compare_states(util.state_of(ref), util.state_of(s), 1e-12, 1e-12, box=n_cell)  # (deposits are atomic sums)
    # upload -> map -> upload (buffers kept), then -> map -> map (released) -> upload again
    for rounds in (1, 2):
        P = [np.ascontiguousarray(x) for x in s.get_particles(0)]
        s.set_particles(0, *P)
        assert np.array_equal(_sorted_cols(s.get_particles(0)), _sorted_cols(P))
        for _ in range(rounds):
            s.map(4, 0.5)
            ref.map(4, 0.5)
    a, b = util.state_of(s), util.state_of(ref)
    errs = util.compare_states(b, a, 1e-12, 1e-12, box=n_cell)
    # a particle outside the box: refused, the species is left empty, and the next upload works
    badp = [x.copy() for x in sets[2]]
    badp[0][3] = n_cell[0] + 0.5
    with pytest.raises(Exception):
        s.set_particles(0, *badp)
    s.set_particles(0, *sets[2])
    assert np.array_equal(_sorted_cols(s.get_particles(0)), _sorted_cols(sets[2]))
    print(engine, errs)


@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("periodic", [(0, 1, 1), (1, 0, 0)])
def test_fused_half_blocks_with_walls(periodic, fuse, interp):
    """Walls (MABC + reflecting particles, the reference's production decks: examples/full/bernstein_main.cpp).
    Theta_B cannot be moved in front of the axis sub-flows there (its MABC blend reads E on a plane the deposits
    reach), so a map2 runs as Theta_E [x y z] Theta_B [z y x] Theta_E with each bracket ONE fused pass; a particle
    that reaches a reflect cell crosses a cell face first and is reflected by the general code (util.hpp:172-186).
    Hot plasma, 14 steps: the walls are reached."""
    n_cell = (18, 12, 10)
    W = 2 if interp == 0 else 1
    E, B = util.rng_fields(n_cell, 93, 0.2)
    parts = np.stack(util.plasma(n_cell, 20, 0.3, 93))
    keep = np.ones(parts.shape[1], bool)
    for d in range(3):  # nothing starts within W + 2 cells of a wall
        if not periodic[d]:
            keep &= (parts[d] >= W + 2) & (parts[d] < n_cell[d] - W - 2)
    parts = [np.ascontiguousarray(t) for t in parts[:, keep]]
    q, m = -1.0 / 20, 100.0 / 20
    o = ora.best_oracle(n_cell, periodic=periodic, interp=interp)
    s = spic().Simulation(n_cell, periodic=periodic, interp=interp)
    s.set_option("fuse", fuse)
    s.set_option("time_kernels", 1)
    for t in (o, s):
        util.load_state(t, E, B, parts, q, m)
    for k in range(14):
        order = 4 if k % 7 == 3 else 2
        o.map(order, 0.5)
        s.map(order, 0.5)
    nmaps2 = 12 + 2 * 3
    errs = util.compare_states(util.state_of(o), util.state_of(s), TOL_STEP * nmaps2, TOL_STEP * nmaps2, box=n_cell)
    kt = s.kernel_times()
    assert kt["axis_block"][1] == (2 * nmaps2 if fuse else 0), kt
    assert kt["theta_axis"][1] == (0 if fuse else 6 * nmaps2), kt
    # the kicks of adjacent Theta_E halves merge on wall boxes too (only the FIELD half is not additive: MABC): one per
    # map2 + the flush in front of the read-back above
    assert kt["push_V_E"][1] == (nmaps2 + 1 if fuse else 2 * nmaps2), kt
    Po = np.stack(o.get_particles())
    assert s.num_particles() == Po.shape[1]
    reached = 0
    for d in range(3):
        if not periodic[d]:
            lo, hi = W + 1, n_cell[d] - 1 - W
            assert np.all((Po[d] >= lo) & (Po[d] < hi)), "a particle entered a reflect cell"
            reached += np.sum(Po[d] - lo < 0.5) + np.sum(hi - Po[d] < 0.5)
    assert reached > 0  # the walls were reached
    print(periodic, fuse, interp, errs, kt)


def test_deferred_half_kick_is_invisible():
    """The last Theta_E half of a fused map is deferred and merged with the first half of the next one
    (Theta_E(s) o Theta_E(t) = Theta_E(s + t)); any observation applies it first.  Three chained Theta_map4 with
    and without deferral, and with an observation between the steps, against the oracle."""
    n_cell = (10, 9, 8)
    E, B = util.rng_fields(n_cell, 91, 0.3)
    parts = util.plasma(n_cell, 12, 0.05, 91)
    q, m = -1.0 / 12, 100.0 / 12
    o = ora.best_oracle(n_cell, interp=0)
    util.load_state(o, E, B, parts, q, m)
    for _ in range(3):
        o.map(4, 0.5)
    ref = util.state_of(o)
    launches = {}
    for mode in ("deferred", "observed", "immediate"):
        s = spic().Simulation(n_cell, interp=0)
        s.set_option("defer_kick", 0 if mode == "immediate" else 1)
        s.set_option("time_kernels", 1)
        util.load_state(s, E, B, parts, q, m)
        for _ in range(3):
            s.map(4, 0.5)
            if mode == "observed":
                s.get_total_energy()
        util.compare_states(ref, util.state_of(s), 3 * TOL_STEP, 3 * TOL_STEP, box=n_cell)
        launches[mode] = s.kernel_times()["push_V_E"][1]
        s.close()
    assert launches == {"deferred": 10, "observed": 12, "immediate": 12}, launches


def test_errors_are_reported():
    sp = spic()
    with pytest.raises(sp.SpicError):
        sp.Simulation((8, 8, 8), interp=0, ng=1)  # ng < interpolation range
    s = sp.Simulation((8, 8, 8), interp=0)
    s.set_uniform_field(1, [0, 0, 0])
    with pytest.raises(sp.SpicError):
        s.add_species(-1.0, 1.0, [9.0], [1.0], [1.0], [0.0], [0.0], [0.0])  # outside the box
    s.add_species(-1.0, 1.0, [4.0], [4.0], [4.0], [5.0], [0.0], [0.0])  # |v dt| = 2.5 cells
    s.G_Theta(0, 0.5)
    with pytest.raises(sp.SpicError) as e:
        s.sync()
    assert e.value.code == -4  # SPIC_ECFL


def test_fp64_probe_runs():
    tf = spic().probe_fp64_tflops(0, 0.2)
    assert 5.0 < tf < 80.0, tf
    tf3 = spic().probe_fp64_tflops(0, 0.2, three_operands=True)  # register-file bound: ~2/3 of the above
    assert 5.0 < tf3 <= tf * 1.02, (tf3, tf)

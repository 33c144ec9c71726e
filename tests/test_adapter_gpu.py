"""The literal drop-in of INTEGRATION.md (include/strugepic_amrex_adapter.hpp), compiled against the AMReX stand-in
together with the reference (oracle/build_oracle.build_ref_adapter -> oracle/_ref/liboracle_adapter_*.so): the
reference's own containers (MultiFab, AoS CParticleContainer) stay on the host, every Theta_map / G_Theta call goes
MultiFab -> C ABI -> GPU -> MultiFab.  Run next to the reference's CPU functions on the same containers' contents.
Mirrors the loop of test/single_particle/main.cpp and test/energy_conservation/main.cpp."""
import numpy as np
import pytest

import oracle as ora
import util

pytestmark = pytest.mark.gpu
TOL_STEP = 1e-11


def _pair(n_cell, periodic, interp):
    import os
    if not (os.path.isfile(ora.adapter_lib_path(interp)) and ora.have_ref(interp)):
        pytest.skip("oracle/_ref/liboracle_adapter_*.so missing (built where /root/reference exists)")
    return (ora.RefOracle(n_cell, periodic=periodic, interp=interp),
            ora.RefOracle(n_cell, periodic=periodic, interp=interp, adapter=True))


@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (0, 1, 1)])
def test_reference_containers_through_the_adapter(periodic, interp):
    n_cell = (16, 9, 6)
    W = 2 if interp == 0 else 1
    o, a = _pair(n_cell, periodic, interp)
    assert a.kind == "adapter"
    E, B = util.rng_fields(n_cell, 21)
    el = util.plasma(n_cell, 6, 0.2, 21, periodic, W)
    io = util.plasma(n_cell, 3, 0.02, 22, periodic, W)
    parts = [np.concatenate([x, y]) for x, y in zip(el, io)]
    q = np.concatenate([np.full(len(el[0]), -1.0 / 6), np.full(len(io[0]), 1.0 / 3)])   # two (q, m) pairs: two species
    m = np.concatenate([np.full(len(el[0]), 100.0 / 6), np.full(len(io[0]), 1836.0 / 3)])
    for t in (o, a):
        t.set_field(0, E)
        t.set_field(1, B)
        t.set_particles(*parts, q, m)
    ops = [("map", 1, 0.5), ("map", 2, 0.5), ("axis", 0, 0.3), ("E", 0.2), ("B", 0.4), ("map", 4, 0.5), ("map", 2, 0.5)]
    for n, op in enumerate(ops, 1):
        util.apply(o, op)
        util.apply(a, op)
        util.compare_states(util.state_of(o), util.state_of(a), 3 * TOL_STEP * n, 3 * TOL_STEP * n, box=n_cell)
        eo, ea = o.energy(), a.energy()   # the reference's get_total_energy on the adapter's containers
        assert abs(sum(ea) - sum(eo)) <= 1e-10 * abs(sum(eo))
    assert a.num_particles() == len(parts[0])


def test_single_particle_deck_through_the_adapter():
    """test/single_particle/main.cpp + cyclotron.input: the first steps and the known answers of SURVEY 8c."""
    Q_E, M_E = -1.60217662e-19, 9.427127615688092e-16
    o, a = _pair((12, 12, 12), (1, 1, 1), 0)
    B = np.zeros((3, 12, 12, 12))
    B[2] = 58.8395
    for t in (o, a):
        t.set_field(0, np.zeros_like(B))
        t.set_field(1, B)
        t.set_particles([6.0], [4.0], [6.0], [0.01], [0.0], [0.01], Q_E, M_E)
    a.map(1, 0.5)
    got = [float(t[0]) for t in a.get_particles()]
    assert got == pytest.approx([6.005, 4.0, 6.005, 0.01, 4.9999997388179764e-05, 0.01], rel=1e-13, abs=1e-18)
    o.map(1, 0.5)
    for _ in range(40):
        o.map(1, 0.5)
        a.map(1, 0.5)
    po, pa = np.array(o.get_particles()).ravel(), np.array(a.get_particles()).ravel()
    assert np.max(np.abs(pa - po)) < 1e-12

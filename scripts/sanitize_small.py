"""Small fused + push + continuation workload for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import strugepic_b200 as spic  # noqa: E402
import util  # noqa: E402

# high counts (k_axis_block, k_push_v_e_v3), low counts (k_axis_block_pair, k_push_v_e_quad), x walls (half-blocks,
# MABC folded into the sweeps), field-only steps (double sweep)
for interp, n_cell, ppc, per in ((0, (8, 6, 5), 70, (1, 1, 1)), (1, (6, 5, 4), 40, (1, 1, 1)), (0, (8, 6, 5), 7, (1, 1, 1)),
                                 (1, (9, 5, 4), 9, (1, 1, 1)), (0, (16, 6, 5), 20, (0, 1, 1))):
    E, B = util.rng_fields(n_cell, 5, 0.3)
    parts = util.plasma(n_cell, ppc, 0.1, 5, per, 2 if interp == 0 else 1)
    s = spic.Simulation(n_cell, periodic=per, interp=interp)
    util.load_state(s, E, B, parts, -1.0 / ppc, 100.0 / ppc)
    s.map(2, 0.5)
    s.map(4, 0.5)
    for k in range(3):
        s.field_only_step(3, 1, 0.1, 0.3, 0.5, k)
    s.get_particles()
    s.gauss_residual()
    print("interp", interp, "energy", s.get_total_energy(), "particles", s.num_particles())
    s.close()

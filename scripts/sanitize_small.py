"""Small fused + push + continuation workload for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import strugepic_b200 as spic  # noqa: E402
import util  # noqa: E402

for interp, n_cell, ppc in ((0, (8, 6, 5), 70), (1, (6, 5, 4), 40)):
    E, B = util.rng_fields(n_cell, 5, 0.3)
    parts = util.plasma(n_cell, ppc, 0.1, 5)
    s = spic.Simulation(n_cell, interp=interp)
    util.load_state(s, E, B, parts, -1.0 / ppc, 100.0 / ppc)
    s.map(2, 0.5)
    s.map(4, 0.5)
    print("interp", interp, "energy", s.get_total_energy(), "particles", s.num_particles())
    s.close()

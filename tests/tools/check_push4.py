"""(test infrastructure: uses the oracle)  Pair-blocked push_V_E (option pushve_kernel = 4) against the oracle and against v3, then its time at 128^3 x 64 ppc."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as ora  # noqa: E402
import strugepic_b200 as spic  # noqa: E402
import util  # noqa: E402

ok = True
for interp, n_cell, ppc, vth in ((0, (9, 7, 6), 7, 0.3), (0, (8, 6, 5), 65, 0.1), (1, (7, 5, 4), 33, 0.2), (0, (5, 3, 2), 1, 0.3)):
    E, B = util.rng_fields(n_cell, 3, 0.4)
    parts = util.plasma(n_cell, ppc, vth, 3)
    o = ora.best_oracle(n_cell, interp=interp)
    s = spic.Simulation(n_cell, interp=interp)
    s.set_option("pushve_kernel", 4)
    for t in (o, s):
        util.load_state(t, E, B, parts, -1.0 / ppc, 100.0 / ppc)
    for k in range(3):  # particles move between the kicks: odd / changing counts per cell
        for t in (o, s):
            util.apply(t, ("E", 0.3))
            util.apply(t, ("map", 2, 0.5))
    try:
        errs = util.compare_states(util.state_of(o), util.state_of(s), 1e-10, 1e-10, box=n_cell)
        print("interp %d %s ppc %d: OK %s" % (interp, n_cell, ppc, {k: float("%.2g" % v) for k, v in errs.items()}))
    except AssertionError as e:
        ok = False
        print("interp %d %s ppc %d: MISMATCH %s" % (interp, n_cell, ppc, str(e)[:300]))
    s.close()
print("PARITY", "OK" if ok else "FAILED")

n = 128
s = spic.Simulation((n, n, n), interp=0)
s.set_uniform_field(0, [0.1, 0.2, 0.3])
s.set_uniform_field(1, [0, 0, 1.0])
s.add_particle_density_uniform(64, 100.0, -1.0, 0.01)
for _ in range(2):
    s.Theta_map2(0.5)
s.set_option("time_kernels", 1)
for variant in (3, 4, 3, 4):
    s.set_option("pushve_kernel", variant)
    s.G_Theta_E(0.1)
    s.kernel_times(reset=True)
    for _ in range(4):
        s.G_Theta_E(0.1)
    kt = s.kernel_times(reset=True)["push_V_E"]
    print("pushve_kernel %d: %.3f ms/launch" % (variant, kt[0] / kt[1]))

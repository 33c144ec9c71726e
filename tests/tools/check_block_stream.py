"""(test infrastructure: uses the oracle)  First thing to run on a GPU in round 2: the cell-spanning fused axis block
(option block_stream = 1, k_axis_block_s) against the oracle on the cases of the fused parity test, then its time
against k_axis_block at 128^3 x 64 ppc.   python tests/tools/check_block_stream.py [cells]"""
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
cases = [(0, (8, 6, 5), 70, 0.05, 4), (0, (8, 8, 8), 40, 0.01, 2), (0, (12, 10, 7), 65, 0.3, 4), (1, (9, 7, 6), 45, 0.1, 4),
         (0, (3, 3, 3), 1, 0.01, 4), (0, (4, 4, 1), 33, 0.2, 2), (0, (16, 2, 2), 100, 0.1, 2)]
for interp, n_cell, ppc, vth, order in cases:
    E, B = util.rng_fields(n_cell, 5, 0.3)
    parts = util.plasma(n_cell, ppc, vth, 5)
    o = ora.best_oracle(n_cell, interp=interp)
    s = spic.Simulation(n_cell, interp=interp)
    s.set_option("block_stream", 1)
    for t in (o, s):
        util.load_state(t, E, B, parts, -1.0 / ppc, 100.0 / ppc)
    try:
        for k in range(3):
            o.map(order, 0.5)
            s.map(order, 0.5)
        errs = util.compare_states(util.state_of(o), util.state_of(s), 1e-10, 1e-10, box=n_cell)
        assert s.num_particles() == len(parts[0])
        g = np.max(np.abs(s.gauss_residual()))
        print("interp %d %s ppc %d vth %g order %d: OK %s" % (interp, n_cell, ppc, vth, order,
                                                             {k: float("%.2g" % v) for k, v in errs.items()}), g)
    except (AssertionError, spic.SpicError) as e:
        ok = False
        print("interp %d %s ppc %d vth %g order %d: FAILED %s" % (interp, n_cell, ppc, vth, order, str(e)[:300]))
    s.close()
print("PARITY", "OK" if ok else "FAILED")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
s = spic.Simulation((n, n, n), interp=0)
s.set_uniform_field(0, [0, 0, 0])
s.set_uniform_field(1, [0, 0, 1.0])
s.add_particle_density_uniform(64, 100.0, -1.0, 0.01)
s.set_option("time_kernels", 1)
for _ in range(3):
    s.Theta_map2(0.5)
for variant in (0, 1, 0, 1):
    s.set_option("block_stream", variant)
    s.Theta_map2(0.5)
    s.kernel_times(reset=True)
    for _ in range(2):
        s.Theta_map2(0.5)
    kt = s.kernel_times(reset=True)["axis_block"]
    print("block_stream %d: %.3f ms/launch" % (variant, kt[0] / kt[1]))
print("energy", s.get_total_energy())

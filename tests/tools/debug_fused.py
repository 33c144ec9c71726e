"""(test infrastructure: uses the oracle)  Development aid: fused axis block vs the oracle on small boxes, errors printed without asserting."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as ora  # noqa: E402
import strugepic_b200 as spic  # noqa: E402
import util  # noqa: E402


def case(n_cell, ppc, vth, interp, order, steps, fuse, dt=0.5, bk=2):
    E, B = util.rng_fields(n_cell, 5, 0.3)
    parts = util.plasma(n_cell, ppc, vth, 5)
    q, m = -1.0 / ppc, 100.0 / ppc
    o = ora.best_oracle(n_cell, interp=interp)
    s = spic.Simulation(n_cell, interp=interp)
    s.set_option("fuse", fuse)
    s.set_option("time_kernels", 1)
    for t in (o, s):
        util.load_state(t, E, B, parts, q, m)
    out = []
    for _ in range(steps):
        o.map(order, dt)
        s.map(order, dt)
        try:
            errs = util.compare_states(util.state_of(o), util.state_of(s), 1e300, 1e300, box=n_cell)
        except AssertionError as e:  # particle count / matching problems
            errs = {"assert": str(e)[:200]}
        out.append(errs)
    kt = s.kernel_times()
    print("n_cell=%s ppc=%d vth=%g interp=%d order=%d fuse=%d n=%d/%d" %
          (n_cell, ppc, vth, interp, order, fuse, s.num_particles(), len(parts[0])))
    for e in out:
        print("    ", {k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in e.items()})
    print("    launches:", {k: v[1] for k, v in kt.items()})
    s.close()


if __name__ == "__main__":
    case((8, 8, 8), 4, 0.0, 0, 2, 2, 1)        # cold: nobody leaves a cell
    case((8, 8, 8), 4, 0.01, 0, 2, 2, 1)       # a few leave
    case((8, 8, 8), 40, 0.01, 0, 2, 2, 1)      # two batches per cell
    case((12, 10, 7), 40, 0.08, 0, 4, 2, 1)    # many leave
    case((12, 10, 7), 40, 0.3, 0, 4, 2, 1)     # queue overflow into the mover list
    case((3, 3, 3), 1, 0.01, 0, 4, 2, 1)       # fewer chunks than warps, sparse cells
    case((12, 10, 7), 40, 0.08, 1, 4, 2, 1)    # PWL
    case((4, 4, 1), 5, 0.2, 0, 2, 3, 1)

"""Generates the golden vectors in this directory from the REFERENCE's own code.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
It drives oracle/_ref (the reference sources compiled unmodified, -O2 -ffp-contract=off)
through seeded schedules and stores inputs + outputs as .npz.  The reference itself ships
no golden vectors; these are the pins for the C port and for the CUDA
library.  Committed together with its outputs so the vectors can be regenerated.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle as ora  # noqa: E402
import util  # noqa: E402

DT = 0.5

CASES = {
    # name: (n_cell, periodic, interp, ppc, v_th, seed, q, m, schedule)
    "p8_subflows": ((8, 6, 5), (1, 1, 1), 0, 2, 0.2, 11, -1.0 / 2, 100.0 / 2,
                    [("E", 0.25), ("axis", 0, 0.25), ("axis", 1, 0.25), ("axis", 2, 0.25), ("B", 0.5)]),
    "p8_map1x3": ((8, 6, 5), (1, 1, 1), 0, 2, 0.2, 12, -1.0 / 2, 100.0 / 2, [("map", 1, DT)] * 3),
    "p8_map2x3": ((7, 8, 6), (1, 1, 1), 0, 3, 0.15, 13, -1.0 / 3, 100.0 / 3, [("map", 2, DT)] * 3),
    "p8_map4x2": ((6, 6, 6), (1, 1, 1), 0, 2, 0.1, 14, -1.0 / 2, 100.0 / 2, [("map", 4, DT)] * 2),
    "pwl_map2x3": ((7, 8, 6), (1, 1, 1), 1, 3, 0.15, 15, -1.0 / 3, 100.0 / 3, [("map", 2, DT)] * 3),
    "pwl_map4x2": ((6, 5, 4), (1, 1, 1), 1, 2, 0.2, 16, -1.0 / 2, 100.0 / 2, [("map", 4, DT)] * 2),
    "p8_wall_map1x6": ((16, 5, 4), (0, 1, 1), 0, 2, 0.3, 17, -1.0 / 2, 100.0 / 2,
                       [("map", 1, DT), ("source", 4, 1, 0.1, 0.3, DT, 0.5), ("map", 1, DT),
                        ("source", 4, 1, 0.1, 0.3, DT, 1.0), ("map", 1, DT), ("map", 1, DT),
                        ("map", 1, DT), ("map", 1, DT)]),
    "pwl_wall_map2x4": ((14, 4, 4), (0, 1, 1), 1, 2, 0.3, 18, -1.0 / 2, 100.0 / 2, [("map", 2, DT)] * 4),
}


# the user-W slot (SPIC_INTERP_USER = 2) with the default file csrc/user_w_default.cu (cubic B-spline, range 2),
# generated from the reference with that file linked over its weak W symbols (oracle/_ref/liboracle_ref_user.so)
USER_CASES = {
    "user_subflows": ((8, 6, 5), (1, 1, 1), 2, 2, 0.2, 31, -1.0 / 2, 100.0 / 2,
                      [("E", 0.25), ("axis", 0, 0.25), ("axis", 1, 0.25), ("axis", 2, 0.25), ("B", 0.5)]),
    "user_map2x3": ((7, 8, 6), (1, 1, 1), 2, 3, 0.15, 32, -1.0 / 3, 100.0 / 3, [("map", 2, DT)] * 3),
    "user_wall_map4x2": ((16, 5, 4), (0, 1, 1), 2, 2, 0.3, 33, -1.0 / 2, 100.0 / 2, [("map", 4, DT)] * 2),
}
W_OF = {0: 2, 1: 1, 2: 2}


def build_case(name):
    n_cell, periodic, interp, ppc, v_th, seed, q, m, schedule = (CASES.get(name) or USER_CASES[name])
    W = W_OF[interp]
    E, B = util.rng_fields(n_cell, seed, amp=0.5)
    parts = util.plasma(n_cell, ppc, v_th, seed, periodic, W)
    return dict(n_cell=n_cell, periodic=periodic, interp=interp, q=q, m=m, schedule=schedule,
                E=E, B=B, parts=parts)


def main():
    ora.ensure_built()
    for name in list(CASES) + list(USER_CASES):
        c = build_case(name)
        o = ora.RefOracle(c["n_cell"], periodic=c["periodic"], interp=c["interp"])
        util.load_state(o, c["E"], c["B"], c["parts"], c["q"], c["m"])
        util.run(o, c["schedule"])
        E, B, P = util.state_of(o)
        en = np.array(o.energy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), E0=c["E"], B0=c["B"], P0=np.stack(c["parts"]),
                            E1=E, B1=B, P1=P, energy1=en)
        print(name, "particles", P.shape[1], "energy", en)
    # W-function table on a fixed argument grid (both variants)
    xs = np.linspace(-2.5, 2.5, 201)
    tab = {}
    for interp, tag in ((0, "p8"), (1, "pwl"), (2, "user")):
        o = ora.RefOracle((4, 4, 4), interp=interp)
        tab[tag + "_W1"] = np.array([o.W1(x) for x in xs])
        tab[tag + "_Wp"] = np.array([o.Wp(x) for x in xs])
        tab[tag + "_I_Wp"] = np.array([o.I_Wp(x, x + 0.37) for x in xs])
        tab[tag + "_I_W1"] = np.array([o.I_W1(x, x + 0.37) for x in xs])
    np.savez_compressed(os.path.join(HERE, "w_tables.npz"), xs=xs, **tab)


if __name__ == "__main__":
    main()

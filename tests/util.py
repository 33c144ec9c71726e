"""Shared helpers of the parity tests: seeded cases, schedules, comparisons."""
import numpy as np

import oracle as ora
from strugepic_b200 import synthetic

P8R2, PWL = 0, 1


def rng_fields(n_cell, seed, amp=1.0):
    nx, ny, nz = n_cell
    r = np.random.default_rng(seed)
    E = amp * r.standard_normal((3, nz, ny, nx))
    B = amp * r.standard_normal((3, nz, ny, nx))
    return E, B


def plasma(n_cell, ppc, v_th, seed, periodic=(1, 1, 1), W=2):
    """Uniform thermal plasma; with a wall in x the W+2 cells next to each wall stay empty
    (particles must never start in the reflect cells)."""
    x, y, z, vx, vy, vz = synthetic.uniform_plasma(n_cell, ppc, v_th, seed)
    if not periodic[0]:
        keep = (x >= W + 2) & (x < n_cell[0] - W - 2)
        x, y, z, vx, vy, vz = (t[keep] for t in (x, y, z, vx, vy, vz))
    return [np.ascontiguousarray(t) for t in (x, y, z, vx, vy, vz)]


def apply(obj, op):
    """Run one schedule entry on an oracle or on a strugepic_b200.Simulation."""
    kind = op[0]
    gpu = hasattr(obj, "G_Theta")
    if kind == "axis":
        (obj.G_Theta if gpu else obj.theta_axis)(op[1], op[2])
    elif kind == "E":
        (obj.G_Theta_E if gpu else obj.theta_E)(op[1])
    elif kind == "B":
        (obj.G_Theta_B if gpu else obj.theta_B)(op[1])
    elif kind == "map":
        obj.map(op[1], op[2])
    elif kind == "source":
        pos, comp, E0, omega, dt, t = op[1:]
        if gpu:
            obj.E_source(pos, comp, E0, omega, dt)(t)
        else:
            obj.source(pos, comp, E0, omega, dt, t)
    else:
        raise ValueError(kind)


def run(obj, schedule):
    for op in schedule:
        apply(obj, op)


def make_oracle(kind, n_cell, periodic, interp, ng=None):
    cls = ora.RefOracle if kind == "ref" else ora.PortOracle
    return cls(n_cell, periodic=periodic, ng=ng, interp=interp)


def load_state(obj, E, B, parts, q, m):
    gpu = hasattr(obj, "G_Theta")
    obj.set_field(0, E)
    obj.set_field(1, B)
    if gpu:
        obj.add_species(q, m, *parts)
    else:
        obj.set_particles(*parts, q, m)


def state_of(obj):
    gpu = hasattr(obj, "G_Theta")
    E = obj.get_field(0)
    B = obj.get_field(1)
    P = obj.get_particles(0) if gpu else obj.get_particles()
    return E, B, np.stack(P)


def match_particles(Pref, Pgot):
    """Permutation that aligns Pgot (6, n) with Pref (6, n): nearest neighbour in phase space."""
    from scipy.spatial import cKDTree
    assert Pref.shape == Pgot.shape, (Pref.shape, Pgot.shape)
    if Pref.shape[1] == 0:
        return np.zeros(0, dtype=np.int64)
    scale = np.array([1, 1, 1, 50, 50, 50.0])[:, None]
    t = cKDTree((Pgot * scale).T)
    d, idx = t.query((Pref * scale).T)
    assert len(np.unique(idx)) == len(idx), "ambiguous particle match"
    return idx


def rel_err(a, b):
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def compare_states(ref, got, tol_field, tol_part, ordered=False, box=None):
    Er, Br, Pr = ref
    Eg, Bg, Pg = got
    if not ordered:
        Pg = Pg[:, match_particles(Pr, Pg)]
    errs = {"E": rel_err(Eg, Er), "B": rel_err(Bg, Br)}
    dx = Pg[:3] - Pr[:3]
    if box is not None:  # a position may sit on either side of a periodic seam
        for d in range(3):
            dx[d] = (dx[d] + box[d] / 2) % box[d] - box[d] / 2
    errs["x"] = float(np.max(np.abs(dx))) if Pr.size else 0.0
    vden = np.max(np.abs(Pr[3:])) if Pr.size else 1.0
    errs["v"] = float(np.max(np.abs(Pg[3:] - Pr[3:])) / (vden if vden > 0 else 1.0)) if Pr.size else 0.0
    assert errs["E"] <= tol_field, errs
    assert errs["B"] <= tol_field, errs
    assert errs["x"] <= tol_part, errs
    assert errs["v"] <= tol_part, errs
    return errs

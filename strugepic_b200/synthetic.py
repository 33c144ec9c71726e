"""Synthetic inputs -- the numpy twin of the device generator.

`uniform_plasma` reproduces, bit for bit, what `spic_load_uniform_plasma` writes on the
device (csrc/particle_math.cuh: synth_particle): for every cell in (k, j, i) order, `ppc`
particles at cell corner + U[0,1)^3 with each velocity component an Irwin-Hall(4) variate
of standard deviation v_th; per-particle charge q/ppc and mass m/ppc as in the reference's
add_particle_density (src/strugepic_util.cpp:267-311).  Only +, -, * are used so that host
and device agree exactly; the reference itself seeds from std::random_device
(util.cpp:269-270) and is not reproducible.
"""
import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _uniform(seed, gid, draw):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (gid * np.uint64(16) + np.uint64(draw) + np.uint64(1)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_plasma(n_cell, ppc, v_th, seed=12345, z_range=None):
    """Returns x, y, z, vx, vy, vz for the cells of the global box `n_cell` (or the z slab
    `z_range=(k0, k1)` of it), in the device loader's order."""
    nx, ny, nz = (int(t) for t in n_cell)
    k0, k1 = (0, nz) if z_range is None else z_range
    kk, jj, ii = np.meshgrid(np.arange(k0, k1), np.arange(ny), np.arange(nx), indexing="ij")
    gcell = ((kk.astype(np.uint64) * np.uint64(ny) + jj.astype(np.uint64)) * np.uint64(nx)
             + ii.astype(np.uint64)).ravel()
    gid = (gcell[:, None] * np.uint64(ppc) + np.arange(ppc, dtype=np.uint64)[None, :]).ravel()
    corner = [np.repeat(a.ravel().astype(np.float64), ppc) for a in (ii, jj, kk)]
    pos = [corner[d] + _uniform(seed, gid, d) for d in range(3)]
    scale = v_th * 1.7320508075688772
    vel = []
    for d in range(3):
        a = _uniform(seed, gid, 3 + 4 * d) + _uniform(seed, gid, 4 + 4 * d)
        b = _uniform(seed, gid, 5 + 4 * d) + _uniform(seed, gid, 6 + 4 * d)
        vel.append(scale * ((a + b) - 2.0))
    return pos[0], pos[1], pos[2], vel[0], vel[1], vel[2]


# ---- density profiles of the reference's loaders (src/strugepic_util.cpp:181-208) ---------------------
def uniform_density(n_cell, i, j, k):
    """uniform_density, util.cpp:202-204."""
    return np.ones(np.broadcast(i, j, k).shape)


def simple_line_density(n_cell, i, j, k):
    """simple_line_density, util.cpp:206-208."""
    return 1.0 * i / 20 + 0.0 * (j + k)


def bernstein_density(n_cell, i, j, k):
    """bernstein_density, util.cpp:181-200: empty within 3 cells of the x walls, Gaussian ramps below
    i = 700 and above i = 1300, 1 in between (nr = 380)."""
    nr = 380
    i = np.asarray(i, dtype=np.int64) + 0 * (np.asarray(j) + np.asarray(k))
    lo, hi = 0, int(n_cell[0]) - 1
    up = np.exp(-((i - (nr + 320)) * (i - (nr + 320))) / (2 * (nr / 3.5) * (nr / 3.5)))
    down = np.exp(-((i - 1300) * (i - 1300)) / 26122.0)
    out = np.where(i < nr + 320, up, np.where(i >= 1300, down, 1.0))
    return np.where((i <= lo + 3) | (i >= hi - 3), 0.0, out)


def density_counts(n_cell, dist_func, ppc_max):
    """Per-cell particle counts `int(dist_func(geom,i,j,k) * ppc_max)` (util.cpp:304) over the GLOBAL box,
    shape (nz, ny, nx), and the RNG stride (>= every count) the device loader keys its counters with."""
    nx, ny, nz = (int(t) for t in n_cell)
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    try:
        dens = np.asarray(dist_func(n_cell, ii, jj, kk), dtype=np.float64)
        assert dens.shape == ii.shape
    except Exception:  # a scalar-only callback
        dens = np.array([[[dist_func(n_cell, i, j, k) for i in range(nx)] for j in range(ny)] for k in range(nz)],
                        dtype=np.float64)
    count = np.trunc(dens * ppc_max).astype(np.int64)  # C++ double -> int conversion truncates
    assert count.min() >= 0
    return count.astype(np.int32), int(max(ppc_max, count.max()))


def density_plasma(n_cell, dist_func, ppc_max, v_th, seed=12345, z_range=None):
    """numpy twin of `spic_load_density_plasma`: x, y, z, vx, vy, vz in the device loader's order."""
    nx, ny, nz = (int(t) for t in n_cell)
    count, stride = density_counts(n_cell, dist_func, ppc_max)
    k0, k1 = (0, nz) if z_range is None else z_range
    kk, jj, ii = np.meshgrid(np.arange(k0, k1), np.arange(ny), np.arange(nx), indexing="ij")
    cnt = count[k0:k1].ravel().astype(np.int64)
    gcell = ((kk.astype(np.uint64) * np.uint64(ny) + jj.astype(np.uint64)) * np.uint64(nx)
             + ii.astype(np.uint64)).ravel()
    start = np.concatenate([[0], np.cumsum(cnt)])
    p = (np.arange(start[-1]) - np.repeat(start[:-1], cnt)).astype(np.uint64)
    gid = np.repeat(gcell, cnt) * np.uint64(stride) + p
    corner = [np.repeat(a.ravel().astype(np.float64), cnt) for a in (ii, jj, kk)]
    pos = [corner[d] + _uniform(seed, gid, d) for d in range(3)]
    scale = v_th * 1.7320508075688772
    vel = []
    for d in range(3):
        a = _uniform(seed, gid, 3 + 4 * d) + _uniform(seed, gid, 4 + 4 * d)
        b = _uniform(seed, gid, 5 + 4 * d) + _uniform(seed, gid, 6 + 4 * d)
        vel.append(scale * ((a + b) - 2.0))
    return pos[0], pos[1], pos[2], vel[0], vel[1], vel[2]

"""strugepic_b200 -- B200-native StrugePIC symplectic PIC step.

Host-side mirror of the reference's operator interface
(include/strugepic_propagators.hpp of MoPHA/strugepic) over the C ABI declared in
include/strugepic_b200.h.  The reference's global update functions keep their names:

    sim.G_Theta(comp, dt)      G_Theta<comp,W>     hpp:347-372
    sim.G_Theta_E(dt)          G_Theta_E<W>        hpp:52-71
    sim.G_Theta_B(dt)          G_Theta_B           cpp:102-113
    sim.Theta_map1/2/4(dt)     Theta_map1/2/4<W>   hpp:548-583
    sim.E_source(...)          E_source functor    hpp:19-32, cpp:13-41
    sim.get_total_energy()     get_total_energy    util.cpp:364-394
    sim.checkpoint(path) / sim.restart(path)   SimulationIO::write(step,true)/read(step)

All compute runs in hand-written sm_100a CUDA kernels inside libstrugepic_b200.so;
importing this module without that library raises ImportError (no fallback).
"""
import ctypes as C

import numpy as np

from . import _lib, synthetic
from ._lib import SpicConfig

P8R2, PWL, USER = 0, 1, 2  # USER: the user-supplied W slot (include/strugepic_user_w.h)
FIELD_E, FIELD_B = 0, 1
MAP4_REFERENCE, MAP4_YOSHIDA = 0, 1
ENGINE_BINNED, ENGINE_DIRECT = 0, 1
X, Y, Z = 0, 1, 2

_dp = C.POINTER(C.c_double)


class SpicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("strugepic_b200 error %d: %s" % (code, msg))
        self.code = code


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def interpolation_range(interp):
    """`interpolation_range` of include/strugepic_w.hpp:16 (USER: what the linked-in user file defines)."""
    return _lib.load().spic_interpolation_range(interp)


def W1(x, interp=P8R2):
    """W1 of include/strugepic_w.hpp:12 (host evaluation of the kernels' inline code)."""
    return _lib.load().spic_W1(interp, x)


def Wp(x, interp=P8R2):
    return _lib.load().spic_Wp(interp, x)


def I_W1(a, b, interp=P8R2):
    return _lib.load().spic_I_W1(interp, a, b)


def I_Wp(a, b, interp=P8R2):
    return _lib.load().spic_I_Wp(interp, a, b)


def probe_fp64_tflops(device=0, seconds=1.0, three_operands=False, immediate=False):
    """Measured DFMA rate (TFLOP/s): chains with two loop-constant operands, with three distinct register
    operands per instruction (what the gathers and the deposition issue; the register file sustains ~2/3), or
    Horner steps with immediate coefficients (the W polynomials: the fastest form)."""
    out = C.c_double(0)
    lib = _lib.load()
    fn = lib.spic_probe_fp64_three_operand_tflops if three_operands else (
        lib.spic_probe_fp64_immediate_tflops if immediate else lib.spic_probe_fp64_tflops)
    rc = fn(device, seconds, C.byref(out))
    if rc:
        raise SpicError(rc, "fp64 probe failed")
    return out.value


def comm_unique_id():
    """ncclUniqueId (128 bytes) to be created on rank 0 and broadcast to every rank."""
    buf = C.create_string_buffer(128)
    rc = _lib.load().spic_comm_unique_id(buf)
    if rc:
        raise SpicError(rc, "ncclGetUniqueId failed (NCCL not loadable?)")
    return buf.raw


def read_plot(path):
    """Reader of `Simulation.write_plot` files -> dict(n_cell, lo, n, E, B, n_density)."""
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:8] == b"SPICPLT1", "not a strugepic_b200 plot file"
    hdr = np.frombuffer(raw, dtype=np.int32, count=14, offset=8)
    n_cell, lo, n = hdr[2:5], hdr[5:8], hdr[8:11]
    cells = int(n[0]) * int(n[1]) * int(n[2])
    data = np.frombuffer(raw, dtype=np.float64, count=7 * cells, offset=8 + 14 * 4)
    shp = (int(n[2]), int(n[1]), int(n[0]))
    return {"n_cell": tuple(int(t) for t in n_cell), "lo": tuple(int(t) for t in lo), "n": tuple(int(t) for t in n),
            "interp": int(hdr[1]), "E": data[:3 * cells].reshape((3,) + shp),
            "B": data[3 * cells:6 * cells].reshape((3,) + shp), "n_density": data[6 * cells:].reshape(shp)}


class Simulation:
    """One rank's brick: fields E, B and the particle species, resident in HBM."""

    def __init__(self, n_cell, periodic=(1, 1, 1), interp=P8R2, ng=0, map4_mode=MAP4_REFERENCE,
                 engine=ENGINE_BINNED, device=0, nranks=1, rank=0):
        self.lib = _lib.load()
        cfg = SpicConfig()
        cfg.n_cell[:] = [int(t) for t in n_cell]
        cfg.periodic[:] = [int(bool(t)) for t in periodic]
        cfg.ng, cfg.interp, cfg.map4_mode, cfg.engine = ng, interp, map4_mode, engine
        cfg.device, cfg.nranks, cfg.rank = device, nranks, rank
        h = C.c_void_p()
        rc = self.lib.spic_create(C.byref(cfg), C.byref(h))
        if rc:
            raise SpicError(rc, self.lib.spic_last_error(None).decode())
        self.h = h
        self.interp = interp
        self.W = interpolation_range(interp)
        lo, n = (C.c_int32 * 3)(), (C.c_int32 * 3)()
        self.lib.spic_local_box(self.h, C.byref(lo), C.byref(n))
        self.lo, self.n = tuple(lo), tuple(n)
        self.n_global = tuple(int(t) for t in n_cell)

    # -- plumbing -----------------------------------------------------------------
    def _ck(self, rc):
        if rc < 0:
            raise SpicError(rc, self.lib.spic_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.spic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self.lib.spic_sync(self.h))

    def set_option(self, name, value):
        self._ck(self.lib.spic_set_option(self.h, name.encode(), float(value)))

    def comm_init(self, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.spic_comm_init(self.h, buf))

    # -- state ----------------------------------------------------------------------
    def field_shape(self):
        return (3, self.n[2], self.n[1], self.n[0])

    def set_uniform_field(self, which, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self._ck(self.lib.spic_set_uniform_field(self.h, which, _p(v)))

    def set_field(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == self.field_shape(), (a.shape, self.field_shape())
        self._ck(self.lib.spic_set_field(self.h, which, _p(a)))

    def get_field(self, which, out=None):
        a = np.empty(self.field_shape()) if out is None else out
        self._ck(self.lib.spic_get_field(self.h, which, _p(a)))
        return a

    def add_species(self, q, m, x, y, z, vx, vy, vz):
        arrs = [np.ascontiguousarray(t, dtype=np.float64) for t in (x, y, z, vx, vy, vz)]
        return self._ck(self.lib.spic_add_species(self.h, q, m, len(arrs[0]), *[_p(t) for t in arrs]))

    def set_particles(self, species, x, y, z, vx, vy, vz):
        arrs = [np.ascontiguousarray(t, dtype=np.float64) for t in (x, y, z, vx, vy, vz)]
        self._ck(self.lib.spic_set_particles(self.h, species, len(arrs[0]), *[_p(t) for t in arrs]))

    def add_particle_density_uniform(self, ppc, m, q, v_th, seed=12345):
        """add_particle_density(geom, P, uniform_density, ppc, m, q, v) -- util.cpp:267-311."""
        return self._ck(self.lib.spic_load_uniform_plasma(self.h, q, m, ppc, v_th, seed))

    def add_particle_density(self, dist_func, ppc_max, m, q, v_th, seed=12345):
        """add_particle_density(geom, P, dist_func, ppc_max, m, q, v) -- util.cpp:267-311 -- with any density
        profile `dist_func(i, j, k)` over GLOBAL cell indices (bernstein_density util.cpp:181-200, ...):
        cell (i,j,k) receives int(dist_func * ppc_max) particles of charge q/ppc_max and mass m/ppc_max."""
        count, stride = synthetic.density_counts(self.n_global, dist_func, ppc_max)
        k0 = self.lo[2]
        local = np.ascontiguousarray(count[k0:k0 + self.n[2]], dtype=np.int32)
        return self._ck(self.lib.spic_load_density_plasma(
            self.h, q, m, ppc_max, stride, v_th, seed, local.ctypes.data_as(C.POINTER(C.c_int32))))

    def num_species(self):
        return self.lib.spic_num_species(self.h)

    def num_particles(self, species=0):
        n = C.c_int64(0)
        self._ck(self.lib.spic_num_particles(self.h, species, C.byref(n)))
        return n.value

    def num_particles_global(self, species=0):
        """TotalNumberOfParticles(): summed over the ranks (collective)."""
        n = C.c_int64(0)
        self._ck(self.lib.spic_num_particles_global(self.h, species, C.byref(n)))
        return n.value

    def add_single_particle(self, pos, vel, m, q):
        """add_single_particle (util.cpp:130-155): every rank adds the species, the rank that owns `pos` holds the
        particle (the reference adds it on grid 0 and Redistributes)."""
        mine = self.lo[2] <= pos[2] < self.lo[2] + self.n[2]
        a = [[t] if mine else [] for t in (*pos, *vel)]
        return self.add_species(q, m, *a)

    def get_particles(self, species=0, out=None):
        n = self.num_particles(species)
        arrs = [np.empty(n) for _ in range(6)] if out is None else out
        self._ck(self.lib.spic_get_particles(self.h, species, *[_p(t) for t in arrs]))
        return arrs

    # -- the reference's operators --------------------------------------------------------
    def G_Theta(self, comp, dt):
        self._ck(self.lib.spic_theta_axis(self.h, comp, dt))

    def G_Theta_E(self, dt):
        self._ck(self.lib.spic_theta_E(self.h, dt))

    def G_Theta_B(self, dt):
        self._ck(self.lib.spic_theta_B(self.h, dt))

    def Theta_map1(self, dt):
        self._ck(self.lib.spic_map(self.h, 1, dt))

    def Theta_map2(self, dt):
        self._ck(self.lib.spic_map(self.h, 2, dt))

    def Theta_map4(self, dt):
        self._ck(self.lib.spic_map(self.h, 4, dt))

    def map(self, order, dt):
        self._ck(self.lib.spic_map(self.h, order, dt))

    def E_source(self, pos, comp, E0, omega, dt):
        """Returns the functor `Source(t)` of the reference (hpp:19-32)."""
        def source(t):
            self._ck(self.lib.spic_source(self.h, pos, comp, E0, omega, dt, t))
        return source

    def field_only_step(self, pos, comp, E0, omega, dt, step):
        self._ck(self.lib.spic_field_only_step(self.h, pos, comp, E0, omega, dt, step))

    def get_total_energy(self):
        out = np.zeros(2)
        self._ck(self.lib.spic_energy(self.h, _p(out)))
        return float(out[0]), float(out[1])

    def gauss_residual(self):
        out = np.empty((self.n[2], self.n[1], self.n[0]))
        self._ck(self.lib.spic_gauss_residual(self.h, _p(out)))
        return out

    def number_density(self):
        """get_particle_number_density<W> (include/strugepic_util.hpp:30-85), [k][j][i] valid cells."""
        out = np.empty((self.n[2], self.n[1], self.n[0]))
        self._ck(self.lib.spic_number_density(self.h, _p(out)))
        return out

    def write_plot(self, path):
        """SimulationIO::write<W>(step) plot output: E, B and the number density of this rank's brick."""
        self._ck(self.lib.spic_plot_write(self.h, str(path).encode()))

    def checkpoint(self, path):
        self._ck(self.lib.spic_checkpoint_write(self.h, str(path).encode()))

    def restart(self, path):
        self._ck(self.lib.spic_checkpoint_read(self.h, str(path).encode()))

    # -- introspection -------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.spic_launch_count(self.h))

    def kernel_times(self, reset=False):
        """{kind: (ms, launches)} for theta_axis / push_V_E / curl / other / axis_block (CUDA events, own stream)."""
        ms, n = (C.c_double * 5)(), (C.c_int64 * 5)()
        self._ck(self.lib.spic_kernel_times(self.h, int(reset), C.byref(ms), C.byref(n)))
        return {k: (ms[i], n[i]) for i, k in enumerate(("theta_axis", "push_V_E", "curl", "other", "axis_block"))}

    def stream(self):
        return self.lib.spic_stream(self.h)

    def kernel_time_ms(self, reset=False):
        ms, n = C.c_double(0), C.c_int64(0)
        self._ck(self.lib.spic_kernel_time_ms(self.h, int(reset), C.byref(ms), C.byref(n)))
        return ms.value, n.value

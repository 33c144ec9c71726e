"""ctypes front-ends of the two CPU oracles (TEST INFRASTRUCTURE ONLY).

`PortOracle`  -> oracle/_build/liboracle_port*.so   (plain-C restatement, always buildable)
`RefOracle`   -> oracle/_ref/liboracle_ref_*.so     (the reference's own sources, unmodified)

Both expose the same methods, named after the reference's operators:
  theta_axis(comp, dt)  G_Theta<comp,W>    include/strugepic_propagators.hpp:347-372
  theta_E(dt)           G_Theta_E<W>       include/strugepic_propagators.hpp:52-71
  theta_B(dt)           G_Theta_B          src/strugepic_propagators.cpp:102-113
  map(order, dt)        Theta_map1/2/4     include/strugepic_propagators.hpp:548-583
  source(...)           E_source           src/strugepic_propagators.cpp:13-41
  energy()              get_total_energy   src/strugepic_util.cpp:364-394

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (strugepic_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P8R2, PWL = 0, 1
_dp = C.POINTER(C.c_double)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i3(v):
    return (C.c_int * 3)(*[int(t) for t in v])


def port_lib_path(fast=False):
    return os.path.join(HERE, "_build", "liboracle_port%s.so" % ("_fast" if fast else ""))


USER = 2  # SPIC_INTERP_USER: the reference with a user-supplied W linked in (build_oracle.build_ref_user)


def ref_lib_path(interp, fast=False):
    if interp == USER:
        return os.path.join(HERE, "_ref", "liboracle_ref_user.so")
    return os.path.join(HERE, "_ref", "liboracle_ref_%s%s.so" % ("p8" if interp == P8R2 else "pwl",
                                                                  "_fast" if fast else ""))


def adapter_lib_path(interp):
    """The reference driven through include/strugepic_amrex_adapter.hpp (build_oracle.build_ref_adapter)."""
    return os.path.join(HERE, "_ref", "liboracle_adapter_%s.so" % ("p8" if interp == P8R2 else "pwl"))


def have_ref(interp=P8R2, fast=False):
    return os.path.isfile(ref_lib_path(interp, fast))


def ensure_built():
    """Build what can be built here (the port always; the reference when /root/reference exists)."""
    import sys
    sys.path.insert(0, HERE)
    try:
        import build_oracle
        build_oracle.build_port()
        build_oracle.build_ref()
        build_oracle.build_ref_user()
        build_oracle.build_ref_adapter()
    finally:
        sys.path.pop(0)


class _Base:
    prefix = ""

    def _f(self, name, restype=None, argtypes=None):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        if argtypes is not None:
            fn.argtypes = argtypes
        return fn

    def _bind_common(self):
        vp, d, i, l = C.c_void_p, C.c_double, C.c_int, C.c_long
        self._destroy = self._f("destroy", None, [vp])
        self._set_field = self._f("set_field", None, [vp, i, _dp])
        self._get_field = self._f("get_field", None, [vp, i, _dp])
        self._set_particles = self._f("set_particles", None, [vp, l] + [_dp] * 8)
        self._num_particles = self._f("num_particles", l, [vp])
        self._get_particles = self._f("get_particles", None, [vp] + [_dp] * 6)
        self._theta_axis = self._f("theta_axis", None, [vp, i, d])
        self._theta_E = self._f("theta_E", None, [vp, d])
        self._theta_B = self._f("theta_B", None, [vp, d])
        self._source = self._f("source", None, [vp, i, i, d, d, d, d])
        self._energy = self._f("energy", None, [vp, _dp])
        self._number_density = self._f("number_density", None, [vp, _dp])

    # -- state -----------------------------------------------------------------
    def set_field(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == (3, self.n[2], self.n[1], self.n[0]), a.shape
        self._set_field(self.h, 0 if which in (0, "E") else 1, _p(a))

    def get_field(self, which):
        a = np.empty((3, self.n[2], self.n[1], self.n[0]))
        self._get_field(self.h, 0 if which in (0, "E") else 1, _p(a))
        return a

    def set_particles(self, x, y, z, vx, vy, vz, q, m):
        n = len(x)
        arrs = [np.ascontiguousarray(t, dtype=np.float64) for t in (x, y, z, vx, vy, vz)]
        qa = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float64), (n,)))
        ma = np.ascontiguousarray(np.broadcast_to(np.asarray(m, dtype=np.float64), (n,)))
        self._set_particles(self.h, n, *[_p(t) for t in arrs], _p(qa), _p(ma))

    def num_particles(self):
        return int(self._num_particles(self.h))

    def get_particles(self):
        n = self.num_particles()
        out = [np.empty(n) for _ in range(6)]
        self._get_particles(self.h, *[_p(t) for t in out])
        return out

    # -- operators ---------------------------------------------------------------
    def theta_axis(self, comp, dt):
        self._theta_axis(self.h, comp, dt)

    def theta_E(self, dt):
        self._theta_E(self.h, dt)

    def theta_B(self, dt):
        self._theta_B(self.h, dt)

    def source(self, pos, comp, E0, omega, dt, t):
        self._source(self.h, pos, comp, E0, omega, dt, t)

    def energy(self):
        out = np.zeros(2)
        self._energy(self.h, _p(out))
        return float(out[0]), float(out[1])

    def number_density(self):
        """get_particle_number_density<W> (include/strugepic_util.hpp:30-85), [k][j][i] valid cells."""
        out = np.empty((self.n[2], self.n[1], self.n[0]))
        self._number_density(self.h, _p(out))
        return out

    def close(self):
        if getattr(self, "h", None):
            self._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PortOracle(_Base):
    prefix = "oport_"
    kind = "port"

    def __init__(self, n_cell, periodic=(1, 1, 1), ng=None, interp=P8R2, fast=False):
        path = port_lib_path(fast)
        if not os.path.isfile(path):
            ensure_built()
        self.lib = C.CDLL(path)
        self.interp = interp
        self.n = tuple(int(t) for t in n_cell)
        self.W = 2 if interp == P8R2 else 1
        self.ng = self.W + 1 if ng is None else ng
        self._bind_common()
        create = self._f("create", C.c_void_p, [C.c_int * 3, C.c_int * 3, C.c_int, C.c_int])
        self.h = C.c_void_p(create(_i3(n_cell), _i3(periodic), self.ng, interp))
        self._map = self._f("map", None, [C.c_void_p, C.c_int, C.c_double, C.c_int])
        self._gauss = self._f("gauss", None, [C.c_void_p, _dp])
        for nm, na in (("W1", 1), ("Wp", 1), ("I_W1", 2), ("I_Wp", 2)):
            setattr(self, "_" + nm, self._f(nm, C.c_double, [C.c_int] + [C.c_double] * na))
        self._cs = self._f("construct_segments", C.c_int, [C.c_double, C.c_double, _dp, C.POINTER(C.c_int)])

    def map(self, order, dt, yoshida=False):
        self._map(self.h, order, dt, 1 if yoshida else 0)

    def gauss(self):
        out = np.empty((self.n[2], self.n[1], self.n[0]))
        self._gauss(self.h, _p(out))
        return out

    def W1(self, x):
        return self._W1(self.interp, x)

    def Wp(self, x):
        return self._Wp(self.interp, x)

    def I_W1(self, a, b):
        return self._I_W1(self.interp, a, b)

    def I_Wp(self, a, b):
        return self._I_Wp(self.interp, a, b)

    def construct_segments(self, x0, x1):
        pts = np.zeros(3)
        idx = (C.c_int * 2)()
        n = self._cs(x0, x1, _p(pts), idx)
        return n, pts, (idx[0], idx[1])


class RefOracle(_Base):
    prefix = "oref_"
    kind = "reference"

    def __init__(self, n_cell, periodic=(1, 1, 1), ng=None, interp=P8R2, fast=False, adapter=False):
        # adapter: the same shell, but maps and sub-flows run through the drop-in adapter -> the GPU library
        path = adapter_lib_path(interp) if adapter else ref_lib_path(interp, fast)
        if adapter:
            self.kind = "adapter"
        if not os.path.isfile(path):
            ensure_built()
        if not os.path.isfile(path):
            raise FileNotFoundError(path + " (reference oracle not built: /root/reference absent)")
        self.lib = C.CDLL(path)
        self.interp = interp
        self.n = tuple(int(t) for t in n_cell)
        self.W = self._f("wrange", C.c_int, [])()
        assert self.W == (2 if interp == P8R2 else 1) or interp == USER
        assert self._f("interpolation_range", C.c_int, [])() == self.W
        self.ng = self.W + 1 if ng is None else ng
        self._bind_common()
        create = self._f("create", C.c_void_p, [C.c_int * 3, C.c_int * 3, C.c_int])
        self.h = C.c_void_p(create(_i3(n_cell), _i3(periodic), self.ng))
        self._map = self._f("map", None, [C.c_void_p, C.c_int, C.c_double])
        for nm, na in (("W1", 1), ("Wp", 1), ("I_W1", 2), ("I_Wp", 2)):
            setattr(self, nm, self._f(nm, C.c_double, [C.c_double] * na))
        self._cs = self._f("construct_segments", C.c_int, [C.c_double, C.c_double, _dp, C.POINTER(C.c_int)])

    def map(self, order, dt, yoshida=False):
        assert not yoshida, "the reference's Theta_map4 always uses alpha=1, beta=-1 (hpp:578)"
        self._map(self.h, order, dt)

    def construct_segments(self, x0, x1):
        pts = np.zeros(3)
        idx = (C.c_int * 2)()
        n = self._cs(x0, x1, _p(pts), idx)
        return n, pts, (idx[0], idx[1])


def best_oracle(*args, **kw):
    """The reference build when present, else the port."""
    interp = kw.get("interp", P8R2)
    if have_ref(interp, kw.get("fast", False)):
        return RefOracle(*args, **kw)
    return PortOracle(*args, **kw)

"""Build recipe of libstrugepic_b200.so (sm_100a only, in-tree).

    python -m strugepic_b200.build [--force] [--verbose]

Every .cu under strugepic_b200/csrc is compiled with
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and linked into strugepic_b200/lib/libstrugepic_b200.so (git-ignored; it travels
to the GPU box with the gpurun snapshot).  NCCL is dlopen'ed at run time, so the
library has no link-time dependency beyond the (static) CUDA runtime.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SPIC_BUILD_TAG=<tag> builds a side-by-side variant (objects in build_<tag>/, lib/libstrugepic_b200_<tag>.so) for
# A/B timing; select it at run time with SPIC_B200_LIBRARY
_TAG = os.environ.get("SPIC_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libstrugepic_b200%s.so" % ("_" + _TAG if _TAG else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "--expt-extended-lambda", "-Xptxas", "-v"] + os.environ.get("SPIC_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs.append(os.path.join(HERE, "..", "include", "strugepic_b200.h"))
    return hs


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


# the user-W slot (include/strugepic_user_w.h): these two translation units are relocatable device code, device-
# linked with each other at the final link; USER_W_DEFAULT is replaced by the user's file with --user-w
RDC = {"user_w.cu", "user_w_default.cu"}
USER_W_DEFAULT = "user_w_default.cu"
# the warp-per-cell kernels over the user-supplied W: these sources are compiled a second time with -DSPIC_USER_W_TU
# (relocatable device code, entry points prefixed user_; csrc/engine.cuh)
USER_W_TWICE = ["particles_fused.cu", "particles_stream.cu"]


def _compile(src, verbose, path=None, obj=None, user_tu=False):
    rdc = src in RDC or path is not None or user_tu
    obj = obj or os.path.join(OBJ, ("userw_tu_" if user_tu else "") + src[:-3] + ".o")
    path = path or os.path.join(CSRC, src)
    if not _stale(obj, [path] + _headers()):
        return obj, ""
    cmd = [NVCC] + ARCH + FLAGS + (["-rdc=true", "-I", os.path.join(HERE, "..", "include")] if rdc else []) + (
        ["-DSPIC_USER_W_TU"] if user_tu else []) + ["-x", "cu", "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
    log = os.path.join(OBJ, ("userw_tu_" if user_tu else "") + src[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stderr)
    return obj, (r.stderr if verbose else "")


def build(force=False, verbose=False, user_w=None, out=None):
    """Builds the library.  user_w = path of a source file defining the five symbols of
    include/strugepic_user_w.h: it takes the place of csrc/user_w_default.cu and the result goes to `out`
    (default lib/libstrugepic_b200_userw.so), leaving the stock library untouched."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = [s for s in _sources() if not (user_w and s == USER_W_DEFAULT)]
    jobs = [(s, False) for s in srcs] + [(s, True) for s in USER_W_TWICE]
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda j: _compile(j[0], verbose, user_tu=j[1]), jobs))
    lib = LIB
    if user_w:
        lib = out or os.path.join(LIBDIR, "libstrugepic_b200_userw.so")
        tag = os.path.splitext(os.path.basename(lib))[0]
        res.append(_compile(os.path.basename(user_w), verbose, path=os.path.abspath(user_w),
                            obj=os.path.join(OBJ, "userw_%s.o" % tag)))
    objs = [o for o, _ in res]
    for _, log in res:
        if log:
            print(log)
    if force or _stale(lib, objs):
        # nvcc device-links the relocatable objects (user_w.o + the W definitions) on the way
        cmd = [NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return lib


DRIVERS = os.path.join(HERE, "..", "drivers")
DRIVER_BIN = os.path.join(DRIVERS, "bin")
CXX = os.environ.get("CXX", "g++")


def build_drivers(names=None):
    """The re-created reference drivers (drivers/*.cpp: plain C++14 over include/strugepic_b200.hpp, linked
    against the in-tree library with an rpath relative to the binary)."""
    os.makedirs(DRIVER_BIN, exist_ok=True)
    inc = os.path.join(HERE, "..", "include")
    deps = [os.path.join(inc, f) for f in os.listdir(inc)] + [os.path.join(DRIVERS, "common.hpp")]
    out = []
    for f in sorted(os.listdir(DRIVERS)):
        if not f.endswith(".cpp") or (names and f[:-4] not in names):
            continue
        exe = os.path.join(DRIVER_BIN, f[:-4])
        src = os.path.join(DRIVERS, f)
        if _stale(exe, [src, LIB] + deps):
            cmd = [CXX, "-std=c++14", "-O2", "-Wall", "-Werror", "-I", inc, src, "-o", exe, "-L", LIBDIR,
                   "-lstrugepic_b200", "-Wl,-rpath,$ORIGIN/../../strugepic_b200/lib", "-pthread"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("driver build failed for %s:\n%s" % (f, r.stderr[-4000:]))
        out.append(exe)
    return out


if __name__ == "__main__":
    def _opt(name):
        return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else None

    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, user_w=_opt("--user-w"),
                out=_opt("--out")))
    if not _opt("--user-w"):
        print("\n".join(build_drivers()))

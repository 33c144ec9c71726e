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
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libstrugepic_b200.so")

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


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _headers()):
        return obj, ""
    cmd = [NVCC] + ARCH + FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
    log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stderr)
    return obj, (r.stderr if verbose else "")


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), _sources()))
    objs = [o for o, _ in res]
    for _, log in res:
        if log:
            print(log)
    if force or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


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
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print("\n".join(build_drivers()))

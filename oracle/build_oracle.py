"""Build recipe for the CPU oracles (TEST INFRASTRUCTURE ONLY).

  oracle/_build/liboracle_port[_fast].so   plain-C restatement (oracle/port/spic_oracle.c)
  oracle/_ref/liboracle_ref_{p8,pwl}[_fast].so
      the reference's own sources, compiled UNMODIFIED where they lie under
      /root/reference against the AMReX stand-in in oracle/amrex_shim/ (the
      reference's Makefile needs an installed AMReX + pkg-config and is not run).

"parity" builds use -O2 -ffp-contract=off (no FMA contraction, so port == reference
bit for bit); "_fast" builds use the reference's own optimisation flags
(Makefile:4: -O3 -mavx2 -march=native) minus -march=native plus -mfma, because the
binaries travel to a GPU box whose host CPU may differ from this container's.

/root/reference does not exist on the GPU box: there the prebuilt oracle/_ref/*.so
(git-ignored, not gpurun-ignored) is used as is.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPIC_REFERENCE_ROOT", "/root/reference")
BUILD = os.path.join(HERE, "_build")
REFOUT = os.path.join(HERE, "_ref")

PARITY = ["-O2", "-ffp-contract=off"]
FAST = ["-O3", "-mavx2", "-mfma"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))


def _newer(out, srcs):
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(s) <= t for s in srcs)


def build_port(force=False):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "port", "spic_oracle.c")
    outs = []
    for name, flags in (("liboracle_port.so", PARITY), ("liboracle_port_fast.so", FAST)):
        out = os.path.join(BUILD, name)
        if force or not _newer(out, [src]):
            _run(["gcc", "-std=c11", "-fPIC", "-shared", "-Wall"] + flags + [src, "-o", out, "-lm"])
        outs.append(out)
    return outs


def reference_present():
    return os.path.isfile(os.path.join(REF, "src", "strugepic_propagators.cpp"))


def build_ref(force=False):
    """Compile the reference in place; returns [] when /root/reference is absent."""
    if not reference_present():
        return []
    os.makedirs(REFOUT, exist_ok=True)
    srcs = [
        os.path.join(REF, "src", "strugepic_propagators.cpp"),
        os.path.join(REF, "src", "strugepic_util.cpp"),
        os.path.join(REF, "src", "interpolation", "interpolation.cpp"),
        os.path.join(HERE, "ref_driver.cpp"),
    ]
    deps = srcs + [os.path.join(HERE, "amrex_shim", "amrex_standin.H")]
    inc = ["-I" + os.path.join(HERE, "amrex_shim"), "-I" + os.path.join(REF, "include"),
           "-I" + os.path.join(REF, "src", "interpolation")]
    outs = []
    for tag, defs in (("p8", ["-DINTERPOLATION_P8R2=1", "-DWRANGE=2"]),
                      ("pwl", ["-DINTERPOLATION_PWL=1", "-DWRANGE=1"])):
        for suffix, flags in (("", PARITY), ("_fast", FAST)):
            out = os.path.join(REFOUT, "liboracle_ref_%s%s.so" % (tag, suffix))
            if force or not _newer(out, deps):
                _run(["g++", "-std=c++14", "-fPIC", "-shared", "-w"] + flags + defs + inc + srcs + ["-o", out])
            outs.append(out)
    return outs


def build_ref_adapter(force=False):
    """The reference + include/strugepic_amrex_adapter.hpp (ref_driver.cpp with -DSPIC_ADAPTER): its maps and
    sub-flows go through the C ABI of ../strugepic_b200/lib/libstrugepic_b200.so (linked with an rpath relative to
    the output) -- the literal drop-in of INTEGRATION.md, compiled against the AMReX stand-in.  Needs the library to
    be built first; [] without /root/reference."""
    lib_dir = os.path.join(HERE, "..", "strugepic_b200", "lib")
    if not reference_present() or not os.path.isfile(os.path.join(lib_dir, "libstrugepic_b200.so")):
        return []
    os.makedirs(REFOUT, exist_ok=True)
    srcs = [
        os.path.join(REF, "src", "strugepic_propagators.cpp"),
        os.path.join(REF, "src", "strugepic_util.cpp"),
        os.path.join(REF, "src", "interpolation", "interpolation.cpp"),
        os.path.join(HERE, "ref_driver.cpp"),
    ]
    pub = os.path.join(HERE, "..", "include")
    deps = srcs + [os.path.join(HERE, "amrex_shim", "amrex_standin.H"), os.path.join(pub, "strugepic_amrex_adapter.hpp"),
                   os.path.join(pub, "strugepic_b200.h")]
    inc = ["-I" + os.path.join(HERE, "amrex_shim"), "-I" + os.path.join(REF, "include"),
           "-I" + os.path.join(REF, "src", "interpolation"), "-I" + pub]
    outs = []
    for tag, defs in (("p8", ["-DINTERPOLATION_P8R2=1", "-DWRANGE=2"]), ("pwl", ["-DINTERPOLATION_PWL=1", "-DWRANGE=1"])):
        out = os.path.join(REFOUT, "liboracle_adapter_%s.so" % tag)
        if force or not _newer(out, deps):
            _run(["g++", "-std=c++14", "-fPIC", "-shared", "-w", "-DSPIC_ADAPTER=1"] + PARITY + defs + inc + srcs +
                 ["-o", out, "-L" + lib_dir, "-lstrugepic_b200", "-Wl,-rpath,$ORIGIN/../../strugepic_b200/lib"])
        outs.append(out)
    return outs


DEFAULT_USER_W = os.path.join(HERE, "..", "strugepic_b200", "csrc", "user_w_default.cu")


def build_ref_user(user_src=DEFAULT_USER_W, wrange=2, tag="user", force=False):
    """The reference with a USER-supplied W linked over its weak defaults (the reference's own override
    mechanism, include/strugepic_w.hpp:7-8): reference sources unmodified + oracle/user_w_adapter.cpp + the
    user's file compiled as plain C++.  -> oracle/_ref/liboracle_ref_<tag>.so; [] without /root/reference."""
    if not reference_present():
        return []
    os.makedirs(REFOUT, exist_ok=True)
    srcs = [
        os.path.join(REF, "src", "strugepic_propagators.cpp"),
        os.path.join(REF, "src", "strugepic_util.cpp"),
        os.path.join(HERE, "ref_driver.cpp"),
        os.path.join(HERE, "user_w_adapter.cpp"),
    ]
    deps = srcs + [user_src, os.path.join(HERE, "amrex_shim", "amrex_standin.H")]
    inc = ["-I" + os.path.join(HERE, "amrex_shim"), "-I" + os.path.join(REF, "include"),
           "-I" + os.path.join(HERE, "..", "include")]
    out = os.path.join(REFOUT, "liboracle_ref_%s.so" % tag)
    if force or not _newer(out, deps):
        _run(["g++", "-std=c++14", "-fPIC", "-shared", "-w"] + PARITY + ["-DWRANGE=%d" % wrange] + inc + srcs +
             ["-x", "c++", user_src, "-o", out])
    return [out]


if __name__ == "__main__":
    force = "--force" in sys.argv
    for o in build_port(force) + build_ref(force) + build_ref_user(force=force) + build_ref_adapter(force):
        print("built", os.path.relpath(o, HERE))

"""The C++ header API (include/strugepic_b200.hpp) mirrors the reference's propagator API: a driver
written like test/single_particle/main.cpp must compile against it (CPU) and reproduce the
cyclotron.input known answers (GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "api_smoke.cpp")
LIBDIR = os.path.join(ROOT, "strugepic_b200", "lib")


def _build(tmp, wrange=2):
    exe = os.path.join(str(tmp), "api_smoke_w%d" % wrange)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-DWRANGE=%d" % wrange,
                           "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", LIBDIR, "-lstrugepic_b200",
                           "-Wl,-rpath," + LIBDIR])
    return exe


@pytest.mark.parametrize("wrange", [1, 2])
def test_reference_style_driver_compiles_and_links(tmp_path, wrange):
    from strugepic_b200 import _lib
    _lib.load()
    assert os.path.isfile(_build(tmp_path, wrange))


@pytest.mark.gpu
def test_reference_style_driver_runs(tmp_path):
    r = subprocess.run([_build(tmp_path, 2)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr

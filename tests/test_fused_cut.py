"""CPU enumeration of the batch cutter of the cell-spanning fused axis block (strugepic_b200/csrc/fused_cut.cuh,
used by k_axis_block_s): every particle handed out once and in order, the rules that keep two stencil buffers and
one set of parked accumulators sufficient, and the batch count it buys at 64 particles per cell."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cutter_enumeration(tmp_path):
    exe = str(tmp_path / "fused_cut_test")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "fused_cut_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    m = re.search(r"64 ppc: ([0-9.]+) batches per cell cell-by-cell, ([0-9.]+) with mixed batches \(ideal ([0-9.]+)\)",
                  r.stdout)
    plain, mixed, ideal = (float(t) for t in m.groups())
    assert plain > 2.4 and mixed < 2.12 and mixed >= ideal

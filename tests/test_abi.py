"""CPU tests: the C-ABI library loads and exports exactly what include/strugepic_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "strugepic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spic_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from strugepic_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_header_compiles_as_c():
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write('#include "strugepic_b200.h"\nint main(void){spic_config c; (void)c; return 0;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", p,
                               "-o", os.path.join(d, "t.o")])


def test_no_silent_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import strugepic_b200 as spic
    with pytest.raises(spic.SpicError) as e:
        spic.Simulation((8, 8, 8))
    assert e.value.code == -3  # SPIC_ENODEV


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "strugepic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_synthetic_generator_statistics():
    import numpy as np
    from strugepic_b200 import synthetic
    x, y, z, vx, vy, vz = synthetic.uniform_plasma((6, 5, 4), 16, 0.01, 12345)
    assert len(x) == 6 * 5 * 4 * 16
    assert np.all((x >= 0) & (x < 6)) and np.all((y >= 0) & (y < 5)) and np.all((z >= 0) & (z < 4))
    # cell-major order, ppc per cell
    assert np.array_equal(np.floor(x[::16]).astype(int)[:6], np.arange(6))
    for v in (vx, vy, vz):
        assert abs(np.std(v) - 0.01) < 5e-4 and abs(np.mean(v)) < 1e-3
    # slab generation is a slice of the global generation
    xs = synthetic.uniform_plasma((6, 5, 4), 16, 0.01, 12345, z_range=(2, 4))
    assert np.array_equal(xs[0], x[len(x) // 2:]) and np.array_equal(xs[5], vz[len(x) // 2:])


def test_density_profile_twin_is_decomposition_independent():
    """synthetic.density_plasma (the numpy twin of spic_load_density_plasma): the z slabs of a decomposed box
    concatenate to the undecomposed draw, a profile of 1 reproduces uniform_plasma, counts follow int(profile * ppc)."""
    import numpy as np
    from strugepic_b200 import synthetic
    n_cell = (9, 4, 6)
    prof = lambda n, i, j, k: 0.25 + 0.125 * ((i + 2 * j + 3 * k) % 7)  # noqa: E731  (up to 1.0)
    whole = np.stack(synthetic.density_plasma(n_cell, prof, 8, 0.05, 42))
    parts = [np.stack(synthetic.density_plasma(n_cell, prof, 8, 0.05, 42, z_range=(k0, k0 + 2))) for k0 in (0, 2, 4)]
    assert np.array_equal(np.concatenate(parts, axis=1), whole)
    counts, stride = synthetic.density_counts(n_cell, prof, 8)
    assert stride == 8 and whole.shape[1] == counts.sum() and counts.min() == 2 and counts.max() == 8
    uni = np.stack(synthetic.density_plasma(n_cell, synthetic.uniform_density, 5, 0.05, 42))
    assert np.array_equal(uni, np.stack(synthetic.uniform_plasma(n_cell, 5, 0.05, 42)))
    over, stride2 = synthetic.density_counts((40, 1, 1), synthetic.simple_line_density, 10)
    assert stride2 == 19 and over[0, 0, 39] == 19  # a profile above 1 widens the RNG key stride


def test_amrex_adapter_compiles_against_the_stand_in(tmp_path):
    """include/strugepic_amrex_adapter.hpp (the literal drop-in of INTEGRATION.md) is C++14 over the AMReX calls the
    reference itself makes: it must compile next to the reference's own headers against oracle/amrex_shim.  Needs the
    reference's headers (this container); on the GPU box the prebuilt oracle/_ref/liboracle_adapter_*.so is what runs
    (tests/test_adapter_gpu.py)."""
    import subprocess
    ref = "/root/reference/include"
    if not os.path.isdir(ref):
        pytest.skip("no /root/reference here")
    src = tmp_path / "use_adapter.cpp"
    src.write_text('#include "strugepic_propagators.hpp"\n#include "strugepic_amrex_adapter.hpp"\n'
                   'void step(amrex::Geometry g, CParticleContainer& P, amrex::MultiFab& E, amrex::MultiFab& B) {\n'
                   '  strugepic_b200_amrex::Theta_map4<2>(g, P, E, B, 0.5);\n'
                   '  strugepic_b200_amrex::G_Theta<0, 1>(g, P, E, B, 0.5);\n'
                   '  strugepic_b200_amrex::G_Theta_E<2>(g, P, E, B, 0.5);\n'
                   '  strugepic_b200_amrex::G_Theta_B<2>(g, P, E, B, 0.5);\n}\n')
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-I", os.path.join(ROOT, "oracle", "amrex_shim"), "-I", ref,
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]

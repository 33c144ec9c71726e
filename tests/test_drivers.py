"""The re-created reference drivers (drivers/*.cpp) and the deck reader.

CPU: the ParmParse-compatible reader against the syntax of the shipped decks; every driver builds, reads its
deck(s) completely and then fails loudly for want of a GPU (no CPU fallback).
GPU: each driver's ENERGY / POS / VEL output against the oracle run on the same deck values and the same
particles (strugepic_b200.synthetic is the numpy twin of the device loaders).
"""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as ora
from strugepic_b200 import build as spbuild
from strugepic_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = os.path.join(ROOT, "drivers", "decks")
INC = os.path.join(ROOT, "include")

DRIVER_OF = {"cyclotron": "single_particle", "cyclotron_borders": "single_particle", "break": "single_particle",
             "reflection": "single_particle", "energy": "energy_conservation", "energy_other": "energy_conservation",
             "energy_64": "energy_conservation", "source_absorb": "field_only", "field_only_256": "field_only",
             "langmuir": "langmuir", "full_256": "langmuir", "bernstein": "bernstein"}


@pytest.fixture(scope="module")
def drivers():
    from strugepic_b200 import _lib
    _lib.load()
    return {os.path.basename(p): p for p in spbuild.build_drivers()}


def run_driver(exe, deck, *overrides, timeout=600, env=None):
    cmd = [exe, os.path.join(DECKS, deck + ".input")] + list(overrides)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


def energies(stdout):
    return np.array([[float(t) for t in ln.split()[1:3]] for ln in stdout.splitlines() if ln.startswith("ENERGY:")])


def test_parmparse_reads_the_reference_deck_syntax(tmp_path):
    exe = str(tmp_path / "parmparse_test")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", INC,
                           os.path.join(ROOT, "tests", "cpp", "parmparse_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    r = subprocess.run([exe, os.path.join(DECKS, "cyclotron.input"), "nsteps=7", "n_cell=4 4 1"],
                       capture_output=True, text=True)
    assert "nsteps=7 n_cell=4,4,1 contains_q=1" in r.stdout, r.stdout


def test_every_deck_names_a_driver():
    decks = sorted(f[:-6] for f in os.listdir(DECKS) if f.endswith(".input"))
    assert decks == sorted(DRIVER_OF)


@pytest.mark.parametrize("deck", sorted(DRIVER_OF))
def test_driver_reads_its_deck_then_requires_a_gpu(drivers, deck):
    """Exit code 3 = deck error, 2 = library error.  Without a GPU the deck must parse completely and the
    run must stop at spic_create with SPIC_ENODEV -- never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    r = run_driver(drivers[DRIVER_OF[deck]], deck, "nsteps=1")
    assert r.returncode == 2, (r.returncode, r.stdout, r.stderr)
    assert "strugepic::Error -3" in r.stderr, r.stderr


def test_driver_reports_a_missing_key(drivers):
    r = run_driver(drivers["bernstein"], "langmuir", "nsteps=1")  # langmuir.input has no x_periodic / source keys
    assert r.returncode == 3 and "x_periodic" in r.stderr, r.stderr


def test_density_profiles_match_the_cpp_header(tmp_path):
    """bernstein_density / simple_line_density: the C++ header's restatement == the numpy twin, cell by cell."""
    src = tmp_path / "dens.cpp"
    src.write_text('#include "strugepic_b200.hpp"\n#include <cstdio>\nusing namespace strugepic;\n'
                   'int main(){Geometry g({1800,2,2},{0,1,1});\n'
                   'for(int i=0;i<1800;++i) std::printf("%d %d\\n",(int)(bernstein_density(g,i,0,0)*4000),'
                   '(int)(simple_line_density(g,i,0,0)*10));}\n')
    exe = str(tmp_path / "dens")
    subprocess.check_call(["g++", "-std=c++14", "-I", INC, str(src), "-o", exe, "-L", spbuild.LIBDIR,
                           "-lstrugepic_b200", "-Wl,-rpath," + spbuild.LIBDIR])
    out = np.array([[int(t) for t in ln.split()] for ln in subprocess.check_output([exe], text=True).splitlines()])
    cb, _ = synthetic.density_counts((1800, 2, 2), synthetic.bernstein_density, 4000)
    cl, _ = synthetic.density_counts((1800, 2, 2), synthetic.simple_line_density, 10)
    assert np.array_equal(out[:, 0], cb[0, 0]) and np.array_equal(out[:, 1], cl[0, 0])
    assert cb[0, 0, :4].sum() == 0 and cb[0, 0, -4:].sum() == 0 and cb.max() == 4000


# ---------------------------------------------------------------------------------------------------------
def _oracle_energy_series(n_cell, periodic, interp, ng, E0, B0, parts, q, m, nsteps, order, dt=0.5, source=None):
    o = ora.best_oracle(n_cell, periodic=periodic, interp=interp, ng=ng)
    nx, ny, nz = n_cell
    o.set_field(0, np.broadcast_to(np.asarray(E0, float)[:, None, None, None], (3, nz, ny, nx)).copy())
    o.set_field(1, np.broadcast_to(np.asarray(B0, float)[:, None, None, None], (3, nz, ny, nx)).copy())
    o.set_particles(*[np.ascontiguousarray(t) for t in parts], q, m)
    out = []
    for step in range(nsteps):
        out.append(o.energy())
        if source is not None:
            o.source(*source, dt, step * dt)
        o.map(order, dt)
    return np.array(out), o


@pytest.mark.gpu
def test_single_particle_driver_cyclotron(drivers):
    r = run_driver(drivers["single_particle"], "cyclotron", "nsteps=3", "precision=17")
    assert r.returncode == 0, r.stderr
    pos = [[float(t) for t in re.findall(r"[-+0-9.e]+", ln[4:])] for ln in r.stdout.splitlines()
           if ln.startswith("POS:")]
    # std::cout default precision (6 digits) on the particle lines, as in the reference's print_Particle_info
    assert pos[0] == [6.0, 4.0, 6.0] and pos[1] == [6.005, 4.0, 6.005] and pos[2][0] == pytest.approx(6.01, abs=1e-5)
    e = energies(r.stdout)
    assert len(e) == 3 and e[0, 1] == pytest.approx(9.427127615688092e-16 * 2e-4 / 2, rel=1e-12)
    assert r.stdout.splitlines()[0].strip() == "1"  # TotalNumberOfParticles


@pytest.mark.gpu
@pytest.mark.parametrize("wrange,order", [(2, 1), (2, 2), (1, 4)])
def test_energy_conservation_driver_tracks_the_oracle(drivers, wrange, order):
    nsteps = 12
    r = run_driver(drivers["energy_conservation"], "energy", "nsteps=%d" % nsteps, "precision=17", "seed=7",
                   "wrange=%d" % wrange, "order=%d" % order)
    assert r.returncode == 0, r.stderr
    n_cell = (4, 4, 1)
    parts = synthetic.uniform_plasma(n_cell, 20, 0.01, 7)
    ref, _ = _oracle_energy_series(n_cell, (1, 1, 1), 0 if wrange == 2 else 1, wrange + 1, (1, 1, 1), (1, 1, 1),
                                   parts, -1.0 / 20, 100.0 / 20, nsteps, order)
    got = energies(r.stdout)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-11 * nsteps
    assert int(r.stdout.splitlines()[0]) == 320


@pytest.mark.gpu
def test_field_only_driver_tracks_the_oracle(drivers, tmp_path):
    nsteps = 150
    r = run_driver(drivers["field_only"], "source_absorb", "n_cell=64 4 4", "nsteps=%d" % nsteps, "omega=0.3",
                   "precision=17", "output_interval=50", "data_folder_name=%s" % (tmp_path / "fo"))
    assert r.returncode == 0, r.stderr
    o = ora.best_oracle((64, 4, 4), periodic=(0, 1, 1), interp=0, ng=3)
    z = np.zeros((3, 4, 4, 64))
    o.set_field(0, z)
    o.set_field(1, z)
    o.set_particles(*[np.zeros(0)] * 6, -1.0, 1.0)
    ref = []
    for step in range(nsteps):
        ref.append(o.energy())
        o.theta_E(0.25)
        o.source(4, 1, 0.1, 0.3, 0.5, 0.5 * step)
        o.theta_B(0.5)
        o.theta_E(0.25)
    ref, got = np.array(ref), energies(r.stdout)
    assert got.shape == ref.shape and got[-1, 0] > 1e-3
    assert np.max(np.abs(got[:, 0] - ref[:, 0])) < 1e-12 * np.max(ref[:, 0])
    import strugepic_b200
    plt = strugepic_b200.read_plot(str(tmp_path / "fo" / "plt100.spic"))
    assert plt["n_cell"] == (64, 4, 4) and np.max(np.abs(plt["E"][1])) > 1e-3


@pytest.mark.gpu
def test_langmuir_driver_tracks_the_oracle_and_restarts(drivers, tmp_path):
    n_cell, ppc, nsteps = (36, 2, 2), 50, 8
    folder = "data_folder_name=%s" % (tmp_path / "lm")
    common = ["n_cell=36 2 2", "ppc=%d" % ppc, "precision=17", "seed=11", "output_interval=-1", folder]
    r = run_driver(drivers["langmuir"], "langmuir", "nsteps=%d" % nsteps, "checkpoint_interval=4", *common)
    assert r.returncode == 0, r.stderr
    q, m, v, Bz = 15.239667683505981, 426461.27834, 0.008668732661691913, 250.39022172065484
    parts = synthetic.uniform_plasma(n_cell, ppc, v, 11)
    ref, _ = _oracle_energy_series(n_cell, (1, 1, 1), 0, 2, (0, 0, 0), (0, 0, Bz), parts, q / ppc, m / ppc, nsteps, 1)
    got = energies(r.stdout)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-11 * nsteps
    # restart from the checkpoint of step 4 (SimulationIO::read): the remaining ENERGY lines are reproduced
    r2 = run_driver(drivers["langmuir"], "langmuir", "nsteps=%d" % nsteps, "start_step=4",
                    "checkpoint_interval=-1", *common)
    assert r2.returncode == 0, r2.stderr
    got2 = energies(r2.stdout)
    assert got2.shape == (4, 2) and np.max(np.abs(got2 - got[4:]) / np.abs(got[4:])) < 1e-13


@pytest.mark.gpu
def test_bernstein_driver_tracks_the_oracle(drivers, tmp_path):
    """Density-profile loader + walls + source through the driver; 1800 x 2 x 2 as shipped, 6 ppc at most."""
    n_cell, ppc, nsteps = (1800, 2, 2), 6, 5
    r = run_driver(drivers["bernstein"], "bernstein", "ppc=%d" % ppc, "nsteps=%d" % nsteps, "precision=17", "seed=3",
                   "output_interval=-1", "checkpoint_interval=-1", "data_folder_name=%s" % (tmp_path / "bs"))
    assert r.returncode == 0, r.stderr
    q, m, v, Bz = 15.239667683505981, 426461.27834, 0.008668732661691913, 250.39022172065484
    parts = synthetic.density_plasma(n_cell, synthetic.bernstein_density, ppc, v, 3)
    assert 0 < len(parts[0]) < 1800 * 4 * ppc
    ref, _ = _oracle_energy_series(n_cell, (0, 1, 1), 0, 2, (0, 0, 0), (0, 0, Bz), parts, q / ppc, m / ppc, nsteps, 1,
                                   source=(4, 1, 1.5185670500857256, 0.014495335501085017))
    got = energies(r.stdout)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-11 * nsteps


@pytest.mark.gpu
def test_langmuir_driver_on_two_gpus_matches_one(drivers, tmp_path):
    """One process per GPU, z slabs, NCCL id exchanged through a file (Simulation::comm_bootstrap): the ENERGY
    lines of a 2-rank run equal those of the 1-rank run of the same deck (fused and reference schedules)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    args = ["n_cell=8 8 16", "ppc=6", "nsteps=6", "precision=17", "seed=5", "order=2", "output_interval=-1",
            "checkpoint_interval=-1", "data_folder_name=%s" % (tmp_path / "mg")]
    one = run_driver(drivers["langmuir"], "langmuir", *args)
    assert one.returncode == 0, one.stderr
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_PORT="29517",
                   SPIC_ID_FILE=str(tmp_path / "nccl_id"))
        procs.append(subprocess.Popen([drivers["langmuir"], os.path.join(DECKS, "langmuir.input")] + args, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1] for o in outs]
    e1, e2 = energies(one.stdout), energies(outs[0][0])
    assert e1.shape == e2.shape == (6, 2) and energies(outs[1][0]).size == 0  # only the IO rank prints
    assert np.max(np.abs(e2 - e1) / np.abs(e1)) < 1e-11


@pytest.mark.gpu
def test_single_particle_driver_on_two_gpus(drivers, tmp_path):
    """add_single_particle under a slab decomposition (the reference adds on grid 0 and Redistributes,
    util.cpp:144-155): every rank adds the species, the owner rank holds the particle, TotalNumberOfParticles is
    global; the orbit equals the 1-rank run while the particle crosses slab faces."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    args = ["nsteps=40", "precision=17", "print_every=1", "vel=0.01 0.0 -0.01"]  # starts at z = 6 = the slab face
    one = run_driver(drivers["single_particle"], "cyclotron", *args)
    assert one.returncode == 0, one.stderr
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_PORT="29518",
                   SPIC_ID_FILE=str(tmp_path / "nccl_id_sp"))
        procs.append(subprocess.Popen([drivers["single_particle"], os.path.join(DECKS, "cyclotron.input")] + args,
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1] for o in outs]
    first = [ln for ln in outs[0][0].splitlines() if not ln.startswith("NCCL version")][0]  # (NCCL's own banner)
    assert first.strip() == "1"  # global count on the IO rank
    e1, e2 = energies(one.stdout), energies(outs[0][0])
    assert e1.shape == e2.shape and np.max(np.abs(e2 - e1) / np.maximum(np.abs(e1), 1e-300)) < 1e-11

    def positions(text):
        return [[float(t) for t in re.findall(r"[-+0-9.e]+", ln[4:])] for ln in text.splitlines() if ln.startswith("POS:")]
    p1 = np.array(positions(one.stdout))
    p2 = np.array(positions(outs[0][0]) + positions(outs[1][0]))
    assert len(p2) == len(p1) == 40  # exactly one rank prints the particle at every step
    assert len(positions(outs[0][0])) > 0 and len(positions(outs[1][0])) > 0  # the particle changed rank
    for row in p1:
        assert np.min(np.max(np.abs(p2 - row), axis=1)) < 1e-9

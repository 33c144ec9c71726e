"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), z-slab decomposition over NCCL.

Every rank owns one slab of the same global problem; after the schedule rank 0 gathers the slabs
and compares the global state with the oracle run on the undecomposed box (the reference's own
multi-box test strategy: max_grid_size < n_cell).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)

import oracle as ora  # noqa: E402
import strugepic_b200 as spic  # noqa: E402
import util  # noqa: E402


def main():
    """python mgpu_worker.py case [case ...]: several cases in one rendezvous (one interpreter start-up per rank)."""
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    ok = True
    for case in (sys.argv[1:] or ["p8"]):
        ok = run_case(case, rank, world, local) and ok
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def run_case(case, rank, world, local):
    interp = {"p8": 0, "pwl": 1, "user": 2}[case.split("_")[0]]  # user: the user-W slot (cubic B-spline pair, range 2)
    opts = case.split("_")[1:]
    # slab thickness: thin = 4 planes (< 2 ng: the halo-sum targets overlap), tall = 12 (> 2 (W + 2): the axis block
    # is split into slab-face planes + interior planes and the exchange overlaps the interior), default 6
    nz = (4 if "thin" in opts else 12 if "tall" in opts else 6) * world
    fuse = 0 if "nofuse" in opts else 1  # fused axis blocks (guard width W + 1) / launch per sub-flow
    n_cell = (12, 10, nz)
    ppc, vth = 6, 0.25
    E, B = util.rng_fields(n_cell, 77, 0.3)
    parts = util.plasma(n_cell, ppc, vth, 77)
    q, m = -1.0 / ppc, 100.0 / ppc

    s = spic.Simulation(n_cell, interp=interp, device=local, nranks=world, rank=rank)
    ids = [spic.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s.comm_init(ids[0])
    s.set_option("fuse", fuse)
    if "serial" in opts:
        s.set_option("overlap", 0)  # exchanges in stream order behind the whole axis block
    k0, k1 = s.lo[2], s.lo[2] + s.n[2]
    assert s.n[2] == nz // world
    s.set_field(0, E[:, k0:k1])
    s.set_field(1, B[:, k0:k1])
    mine = (parts[2] >= k0) & (parts[2] < k1)
    s.add_species(q, m, *[t[mine] for t in parts])

    schedule = [("map", 1, 0.5), ("map", 2, 0.5), ("map", 2, 0.5), ("map", 4, 0.5),
                ("E", 0.2), ("axis", 2, -0.7), ("axis", 2, 0.7), ("B", 0.3)]
    util.run(s, schedule)
    s.sync()
    en = s.get_total_energy()
    gauss = s.gauss_residual()  # collective: rho and E guards cross the slab faces
    Es, Bs, Ps = util.state_of(s)
    assert np.all((Ps[2] >= k0) & (Ps[2] < k1)), "a particle sits outside its rank's slab"
    gathered = [None] * world
    dist.gather_object((Es, Bs, Ps, en, gauss), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        Eg = np.concatenate([g[0] for g in gathered], axis=1)
        Bg = np.concatenate([g[1] for g in gathered], axis=1)
        Pg = np.concatenate([g[2] for g in gathered], axis=1)
        o = ora.best_oracle(n_cell, interp=interp)
        util.load_state(o, E, B, parts, q, m)
        util.run(o, schedule)
        try:
            assert Pg.shape[1] == len(parts[0]), "particle count changed: %d != %d" % (Pg.shape[1], len(parts[0]))
            moved = sum(int(g[2].shape[1]) for g in gathered)
            errs = util.compare_states(util.state_of(o), (Eg, Bg, Pg), 1e-10, 1e-10, box=n_cell)
            eo = o.energy()
            for g in gathered:  # every rank holds the allreduced energy
                assert np.allclose(g[3], eo, rtol=1e-10), (g[3], eo)
            # discrete Gauss residual of the decomposed run == the C port's on the oracle's final state
            if interp == 2:  # (the port knows the two shipped variants only)
                raise StopIteration
            po = ora.PortOracle(n_cell, interp=interp)
            Eo, Bo, Po = util.state_of(o)
            util.load_state(po, Eo, Bo, list(Po), q, m)
            go = po.gauss()
            gg = np.concatenate([g[4] for g in gathered], axis=0)
            gerr = float(np.max(np.abs(gg - go)) / np.max(np.abs(go)))
            assert gerr < 1e-10, gerr
            errs["gauss"] = gerr
        except StopIteration:
            pass
        except AssertionError as e:
            print("MULTI-GPU PARITY FAILED:", e)
            ok = False
    if rank == 0 and ok:
        print("multi-gpu parity ok case=%s world=%d particles=%d errs=%s" % (case, world, moved, errs))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    s.close()
    return flag[0]


if __name__ == "__main__":
    main()

"""CPU tests of the N > 1 host logic: world_size-2 gloo processes (no GPU)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))  # decomp.py: the host-side twin of the slab logic


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import decomp
    from strugepic_b200 import synthetic
    n_cell, ppc = (6, 5, 8), 4
    # unique-id plumbing exactly as bench.py / tests/mgpu_worker.py do it (a fake 128-byte id here)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    k0, k1 = decomp.slab_range(n_cell[2], world, rank)
    mine = synthetic.uniform_plasma(n_cell, ppc, 0.01, 99, z_range=(k0, k1))
    assert np.all(decomp.owner_of(mine[2], n_cell[2], world) == rank)
    prev, nxt = decomp.ring_neighbours(rank, world)
    # ring exchange of one "face plane" with gloo: what I send to next must arrive from prev
    import torch
    send = torch.full((4,), float(rank))
    recv = torch.empty(4)
    reqs = [dist.isend(send, nxt), dist.irecv(recv, prev)]
    for r in reqs:
        r.wait()
    assert float(recv[0]) == float(prev)
    gathered = [None] * world
    dist.all_gather_object(gathered, [t.copy() for t in mine])
    if rank == 0:
        whole = synthetic.uniform_plasma(n_cell, ppc, 0.01, 99)
        for d in range(6):
            assert np.array_equal(np.concatenate([g[d] for g in gathered]), whole[d])
        split = decomp.split_particles(whole, n_cell[2], world)
        for r in range(world):
            for d in range(6):
                assert np.array_equal(split[r][d], gathered[r][d])
        out.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_slab_partition_and_ring_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29531
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def _planner_worker(rank, world, port, out):
    """Every rank runs the migration-message protocol of csrc/comm.cu with gloo in place of NCCL: capacities are
    planned from local history only, the payload carries its count in the header, and the receiver must have planned
    exactly the capacity the sender used -- at every exchange, for counts that jump around."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import decomp
    import torch
    prev, nxt = decomp.ring_neighbours(rank, world)
    plan = decomp.MigrationPlanner(cap=200000)
    rng = np.random.default_rng(100 + rank)
    ok = True
    for k in range(12):
        m_lo, m_hi, m_rp, m_rn = plan.plan()
        n_lo, n_hi = (int(rng.integers(0, 30000)) for _ in range(2))  # leavers through my low / high face
        assert n_lo <= m_lo and n_hi <= m_hi, "4 x head-room over the count two exchanges ago"
        send_lo = torch.zeros(1 + m_lo, dtype=torch.float64)
        send_hi = torch.zeros(1 + m_hi, dtype=torch.float64)
        send_lo[0], send_hi[0] = n_lo, n_hi
        send_lo[1:1 + n_lo] = rank + 0.25
        send_hi[1:1 + n_hi] = rank + 0.75
        recv_prev = torch.empty(1 + m_rp, dtype=torch.float64)  # what prev sent through ITS high face
        recv_next = torch.empty(1 + m_rn, dtype=torch.float64)  # what next sent through ITS low face
        # (gloo checks the sizes of a matched send / recv pair: a mismatch in the planned capacities fails here)
        reqs = [dist.isend(send_lo, prev, tag=2 * k), dist.irecv(recv_next, nxt, tag=2 * k),
                dist.isend(send_hi, nxt, tag=2 * k + 1), dist.irecv(recv_prev, prev, tag=2 * k + 1)]
        for r in reqs:
            r.wait()
        c_prev, c_next = int(recv_prev[0]), int(recv_next[0])
        ok = ok and bool(torch.all(recv_prev[1:1 + c_prev] == prev + 0.75)) and bool(torch.all(recv_next[1:1 + c_next] == nxt + 0.25))
        plan.record(n_lo, n_hi, c_prev, c_next)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        out.put("ok" if all(flags) else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_migration_messages_agree_without_a_count_exchange(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_planner_worker, args=(r, world, 29540 + world, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_decomp_helpers():
    import decomp
    assert decomp.slab_range(16, 4, 2) == (8, 12)
    assert decomp.ring_neighbours(0, 4) == (3, 1) and decomp.ring_neighbours(3, 4) == (2, 0)
    g = decomp.guard_planes(8, 2)
    assert g["send_to_next"] == (6, 8) and g["recv_from_prev"] == (-2, 0)
    a = np.arange(3 * 2 * 2 * 2, dtype=float).reshape(3, 2, 2, 2)
    assert decomp.assemble_field([a, a + 100]).shape == (3, 4, 2, 2)
    import pytest
    with pytest.raises(ValueError):
        decomp.slab_range(10, 4, 0)


def test_reference_arm_other_ranks_exit_without_work():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""

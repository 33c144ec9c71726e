"""Host-side helpers of the z-slab decomposition (one brick per GPU, periodic ring).

The reference decomposes the grid into AMReX boxes (`ba.maxSize(max_grid_size)`,
test/single_particle/main.cpp:104) and lets `DistributionMapping` place them on MPI ranks; here
every rank owns one slab of whole z-planes, so each rank has exactly two neighbours and every
face bundle is contiguous in the `[comp][k][j][i]` field layout (csrc/comm.cu).
"""
import numpy as np


def slab_range(nz, nranks, rank):
    """(k0, k1) of this rank's slab; nz must divide evenly (spic_create enforces the same)."""
    if nz % nranks:
        raise ValueError("n_cell[2] = %d is not divisible by nranks = %d" % (nz, nranks))
    per = nz // nranks
    return rank * per, (rank + 1) * per


def ring_neighbours(rank, nranks):
    """(prev, next) on the periodic ring: low-face and high-face neighbour."""
    return (rank - 1) % nranks, (rank + 1) % nranks


def owner_of(z, nz, nranks):
    """Rank owning global z coordinate(s) z in [0, nz)."""
    per = nz // nranks
    return np.minimum(np.floor(np.asarray(z) / per).astype(np.int64), nranks - 1)


def split_particles(parts, nz, nranks):
    """Partition SoA arrays (x, y, z, vx, vy, vz) by owning slab; returns one tuple per rank."""
    own = owner_of(parts[2], nz, nranks)
    return [tuple(np.ascontiguousarray(t[own == r]) for t in parts) for r in range(nranks)]


def assemble_field(slabs):
    """Concatenate per-rank [3][nz_local][ny][nx] arrays (rank order) into the global field."""
    return np.concatenate(list(slabs), axis=1)


def guard_planes(nz_local, ng):
    """Index ranges (local k, guard cells negative) exchanged with the neighbours, as csrc/comm.cu does:
    fill: owner planes `send_*` -> the neighbour's guard planes `recv_*`; sum: the reverse direction."""
    return {
        "send_to_next": (nz_local - ng, nz_local), "recv_from_prev": (-ng, 0),
        "send_to_prev": (0, ng), "recv_from_next": (nz_local, nz_local + ng),
    }


# ---- migration messages without a count read-back (csrc/comm.cu: plan_messages) ---------------------------------------
def message_capacity(count_two_exchanges_ago, cap):
    """Capacity M (particles) of a migration message: a function of the count that crossed the same face two
    exchanges ago -- a number BOTH ends of the pair hold (the sender counted it, the receiver read it from the
    message header), so they agree on M without talking to each other; None = no history yet: the whole buffer."""
    if count_two_exchanges_ago is None:
        return cap
    return min(cap, 4 * int(count_two_exchanges_ago) + 65536)


class MigrationPlanner:
    """One rank's view of the protocol: history of {sent low, sent high, received from prev, received from next}."""

    def __init__(self, cap):
        self.cap = cap
        self.hist = []  # one (sent_lo, sent_hi, recv_prev, recv_next) per completed exchange

    def plan(self):
        """(M_send_lo, M_send_hi, M_recv_prev, M_recv_next) of the next exchange."""
        h = self.hist[-2] if len(self.hist) >= 2 else (None,) * 4
        return tuple(message_capacity(c, self.cap) for c in h)

    def record(self, sent_lo, sent_hi, recv_prev, recv_next):
        self.hist.append((sent_lo, sent_hi, recv_prev, recv_next))

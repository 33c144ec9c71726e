"""CPU enumeration of the batch composition of the low-count kernels (host twins of the device logic; the CUDA kernels
themselves are covered on the GPU by tests/test_parity_gpu.py::test_fused_block_batch_shapes and the low-ppc goldens).

  k_axis_block_pair  (csrc/particles_fused.cu):  a chunk of 8 cells; two consecutive cells with <= 16 particles each
                     share a batch, lanes dealt so that a lane's record slot (= its lane) falls into its cell's half of
                     the deposition's particle subsets; any other cell runs alone in batches of 32.
  k_push_v_e_quad    (csrc/particles_stream.cu):  a chunk of 8 cells; up to four consecutive cells per batch while they
                     fit 32 lanes; a cell with more than 32 particles takes batches of its own.
Invariants: every particle of every cell is handed out exactly once, in order inside its cell; a batch never holds more
than 32 particles; in a pair the two cells' particles sit in disjoint halves of the particle subsets.
"""
import numpy as np
import pytest

CHUNK = 8


def pair_batches(cnt, nsub):
    """Batches of k_axis_block_pair for one chunk (particles_fused.cu:671-677: `pair`, `isB`, `slot`):
    list of dict(lanes -> (cell, particle index) or None)."""
    half = nsub // 2
    out, ci, off = [], 0, 0
    while ci < CHUNK:
        cnt_a = cnt[ci]
        if cnt_a == 0:
            ci += 1
            continue
        pair = off == 0 and cnt_a <= 16 and ci + 1 < CHUNK and 0 < cnt[ci + 1] <= 16
        lanes = []
        for lane in range(32):
            if pair:
                is_b = (lane % nsub) >= half
                slot = (lane // nsub) * half + lane % half
                cell = ci + (1 if is_b else 0)
                lanes.append((cell, slot) if slot < cnt[cell] else None)
            else:
                slot = off + lane
                lanes.append((ci, slot) if slot < cnt_a else None)
        out.append({"pair": pair, "lanes": lanes})
        if pair or off + 32 >= cnt_a:
            ci += 2 if pair else 1
            off = 0
        else:
            off += 32
    return out


def quad_batches(cnt):
    """Batches of k_push_v_e_quad for one chunk (particles_stream.cu:335-342): list of lists of (cell, particle index)."""
    out, ci, off = [], 0, 0
    while ci < CHUNK:
        tot = min(32, cnt[ci] - off)
        whole0 = off + tot >= cnt[ci]
        nc, bounds = 1, [tot]
        if whole0:
            while nc < 4 and ci + nc < CHUNK and tot + cnt[ci + nc] <= 32:
                tot += cnt[ci + nc]
                bounds.append(tot)
                nc += 1
        batch = []
        for lane in range(tot):
            k = next(i for i, b in enumerate(bounds) if lane < b)
            batch.append((ci + k, off + lane if k == 0 else lane - bounds[k - 1]))
        out.append(batch)
        if whole0:
            ci, off = ci + nc, 0
        else:
            off += 32
    return out


def _expected(cnt):
    return [(c, p) for c in range(CHUNK) for p in range(cnt[c])]


@pytest.mark.parametrize("nsub", [4, 16])  # W8: 4 particle subsets; PWL: 16
def test_pair_batches_hand_out_every_particle_once(nsub):
    rng = np.random.default_rng(3)
    shapes = [rng.poisson(lam, CHUNK) for lam in (1, 4, 8, 8, 12, 16, 20, 40, 70) for _ in range(40)]
    shapes += [np.array(t) for t in ([0] * 8, [16] * 8, [17, 16, 16, 0, 0, 1, 33, 16], [16, 0, 16, 16, 0, 0, 0, 16],
                                     [64, 65, 1, 1, 1, 32, 31, 16], [1, 2, 3, 4, 5, 6, 7, 8])]
    pairs = singles = 0
    for cnt in shapes:
        seen = []
        for b in pair_batches(cnt, nsub):
            live = [t for t in b["lanes"] if t is not None]
            assert len(live) <= 32
            seen += live
            if b["pair"]:
                pairs += 1
                cells = sorted({t[0] for t in live})
                assert len(cells) <= 2 and all(cnt[c] <= 16 for c in cells)
                for lane, t in enumerate(b["lanes"]):  # record slot = lane: subset = lane % nsub, half = the cell
                    if t is not None:
                        assert ((lane % nsub) >= nsub // 2) == (t[0] == cells[-1] and len(cells) == 2 or
                                                                (len(cells) == 1 and (lane % nsub) >= nsub // 2))
            else:
                singles += 1
                assert len({t[0] for t in live}) == 1
        assert sorted(seen) == _expected(cnt), cnt
    assert pairs > 100 and singles > 100


def test_quad_batches_hand_out_every_particle_once():
    rng = np.random.default_rng(4)
    shapes = [rng.poisson(lam, CHUNK) for lam in (1, 4, 8, 8, 12, 16, 33, 70) for _ in range(40)]
    shapes += [np.array(t) for t in ([0] * 8, [8] * 8, [32, 0, 0, 32, 1, 31, 0, 0], [33, 31, 1, 0, 0, 0, 64, 0])]
    lanes_used = batches = 0
    for cnt in shapes:
        seen = []
        for b in quad_batches(cnt):
            assert len(b) <= 32 and len({t[0] for t in b}) <= 4
            seen += b
            lanes_used += len(b)
            batches += 1
        assert seen == _expected(cnt), cnt  # in cell order, in order inside every cell
    # at 8 particles per cell a batch is full (4 x 8): the reason the kernel exists
    full = quad_batches(np.full(CHUNK, 8))
    assert len(full) == 2 and all(len(b) == 32 for b in full)
    assert lanes_used / batches > 12

"""Host-side model of mdct_binmajor_kernel / imdct_binmajor_kernel (zaf-python_b200/csrc/mdct_binmajor.cu): the index
arithmetic of the two kernels replayed in Python, no GPU.

MDCT (store phase): which thread writes which (row, frame) of the reference's C-order (M, nt) matrix.
  * every element is written exactly once, from the frame of the same index;
  * a frame is only read from the ring while it is still there (at most 7 frames behind the current tile of 32);
  * away from run boundaries every 32-frame store run starts on a 32-byte sector boundary, whatever nt mod 8 and the
    4-byte phase of the result pointer are.
IMDCT (TDAC hand-off): warp w of a tile transforms frames j0 + 2w and j0 + 2w + 1; hop-block h must be
  second half of frame h - 1 + first half of frame h, every block 1 .. nt - 1 exactly once, and a parked half must still
  be in its slot (ring of 33) when it is read.
"""
import numpy as np
import pytest

F, WARPS = 32, 16


def replay_mdct_store(m, nt, phase0, runs_per_clip):
    slots = F + 7
    tiles_per_clip = -(-nt // F)
    tiles_per_run = -(-tiles_per_clip // runs_per_clip)
    runs_per_clip = -(-tiles_per_clip // tiles_per_run)
    written = np.zeros((m, nt), np.int32)
    unaligned_interior = 0
    for run in range(runs_per_clip):
        t0 = run * tiles_per_run
        t1 = min(t0 + tiles_per_run, tiles_per_clip)
        jlo, jhi = t0 * F, min(nt, t1 * F)
        ring = {}
        for t in range(t0, t1 + 1):
            j0 = t * F
            if t < t1:
                for warp in range(WARPS):
                    for h in range(2):
                        j = j0 + 2 * warp + h
                        if j < nt:
                            ring[j % slots] = j
            for warp in range(WARPS):
                s_row = (phase0 + (warp & 7) * (nt & 7)) & 7
                for lane in range(32):
                    ja = j0 - s_row + lane
                    if jlo <= ja < jhi:
                        assert ring[ja % slots] == ja, "frame no longer in the ring"
                        for i in range(m // WARPS):
                            row = warp + WARPS * i
                            written[row, ja] += 1
                            if lane == 0 and ja > jlo and (phase0 + row * nt + ja) % 8:
                                unaligned_interior += 1
    return written, unaligned_interior


@pytest.mark.parametrize("m", [64, 128])
@pytest.mark.parametrize("nt", [1, 2, 31, 32, 33, 39, 64, 71, 101])
def test_mdct_store_every_element_once_and_sector_aligned(m, nt):
    for phase0 in range(8):
        for runs in (1, 2, 3):
            written, unaligned = replay_mdct_store(m, nt, phase0, runs)
            assert np.all(written == 1), (m, nt, phase0, runs)
            assert unaligned == 0, (m, nt, phase0, runs)


def replay_imdct_blocks(nt):
    """-> {hop-block: (frame whose second half, frame whose first half)} as the kernel combines them."""
    slots = F + 1
    blocks = {}
    tiles = -(-nt // F)
    parked_second = {}   # slot -> frame whose windowed second half sits there
    parked_first = {}    # slot -> frame whose first half sits there
    for t in range(tiles):
        j0 = t * F
        loaded = {(j0 + w) % slots: j0 + w for w in range(F) if j0 + w < nt}
        # the load phase must not overwrite the carry of the previous tile's last frame
        assert (j0 - 1) % slots not in loaded
        for s in loaded:               # a slot that receives a new spectrum loses what was parked in it
            parked_second.pop(s, None)
            parked_first.pop(s, None)
        for warp in range(WARPS):
            ja = j0 + 2 * warp
            if ja >= nt:
                continue
            parked_first[ja % slots] = ja          # h = 0: first half parked, second half carried in registers
            if ja + 1 < nt:                         # h = 1: block ja + 1 = carry (frame ja) + first half of frame ja + 1
                assert ja + 1 not in blocks
                blocks[ja + 1] = (ja, ja + 1)
                parked_second[(ja + 1) % slots] = ja + 1
        for warp in range(WARPS):                  # combine phase, after the barrier
            ja = j0 + 2 * warp
            if 1 <= ja < nt:
                assert parked_first.get(ja % slots) == ja, "first half no longer parked"
                assert parked_second.get((ja - 1) % slots) == ja - 1, "second half no longer parked"
                assert ja not in blocks
                blocks[ja] = (ja - 1, ja)
    return blocks


@pytest.mark.parametrize("nt", [2, 3, 31, 32, 33, 34, 64, 65, 66, 97, 1293])
def test_imdct_every_hop_block_once_in_reference_order(nt):
    blocks = replay_imdct_blocks(nt)
    assert sorted(blocks) == list(range(1, nt))          # hop-blocks 1 .. nt-1 (zaf.py:1182 trims block 0)
    for h, (a, b) in blocks.items():
        assert (a, b) == (h - 1, h)                      # frame h-1's second half first, then frame h's first half

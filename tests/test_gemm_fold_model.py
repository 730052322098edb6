"""Host-side model of the converter warps of gemm3xtf32_pair_kernel<FOLD> (zaf-python_b200/csrc/gemm_tc.cu): the index
arithmetic that turns the two RAW tiles TMA drops into a pipeline stage -- x[row][32 kb .. + 32) in the A_hi slot and the
mirrored block x[row][n - 32 (kb + 1) .. + 32) in the A_lo slot, both in the 128-byte swizzle -- into the TF32 hi / lo
operand tiles of s = x[m] + x[n-1-m] or d = x[m] - x[n-1-m], in place.  Replayed in NumPy it pins, without a GPU:
  * the swizzle: element (row, col) of a 128 x 32 fp32 tile sits in 16-byte chunk (col / 4) xor (row mod 8) of its row;
  * the pairing: fold index m = 32 kb + i meets element 31 - i of the mirrored block, i.e. chunk j meets chunk 7 - j reversed,
    and one thread owns chunks j and 7 - j of a row in BOTH slots, so reading all four before writing makes it in place;
  * the coverage: converter warps x iterations x lanes touch every chunk of the tile exactly once;
  * the split: hi + lo reproduces the fold value to 2^-21 and both halves are TF32-representable (13 zero mantissa bits),
    with the integer rounding the kernel uses equal to the closed form of cvt.rna.tf32.f32 (ties away from zero).
"""
import numpy as np
import pytest

ROWS, BK = 128, 32


def swz(row, chunk):
    return chunk ^ (row & 7)


def tma_tile(block):
    """A (128, 32) fp32 block as TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B: 8 chunks of 4 floats per row."""
    tile = np.zeros((ROWS, 8, 4), np.float32)
    for r in range(ROWS):
        for j in range(8):
            tile[r, swz(r, j)] = block[r, 4 * j:4 * j + 4]
    return tile


def rna_tf32(v):
    """cvt.rna.tf32.f32 for finite values, as two integer operations on the bit pattern (gemm_tc.cu: tf32_split)."""
    u = np.asarray(v, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def convert_in_place(hi_slot, lo_slot, sign, warps=2):
    """The converter loop of the kernel: cw = converter warp, it = iteration, lane -> (rl, jp)."""
    touched = np.zeros((ROWS, 8), np.int32)
    for cw in range(warps):
        for it in range(ROWS // (8 * warps)):
            for lane in range(32):
                rl, jp = lane & 7, lane >> 3
                row = (ROWS // warps) * cw + 8 * it + rl
                oa, ob = swz(row, jp), swz(row, 7 - jp)
                f0, f1 = hi_slot[row, oa].copy(), hi_slot[row, ob].copy()
                m0, m1 = lo_slot[row, oa].copy(), lo_slot[row, ob].copy()
                v0 = (f0 + np.float32(sign) * m1[::-1]).astype(np.float32)   # chunk jp pairs with mirrored chunk 7 - jp, reversed
                v1 = (f1 + np.float32(sign) * m0[::-1]).astype(np.float32)
                for o, v in ((oa, v0), (ob, v1)):
                    h = rna_tf32(v)
                    hi_slot[row, o] = h
                    lo_slot[row, o] = rna_tf32((v - h).astype(np.float32))
                    touched[row, o] += 1
    return touched


@pytest.mark.parametrize("n,kb", [(1024, 0), (1024, 15), (1000, 3), (1000, 15), (64, 0), (16, 0)])
@pytest.mark.parametrize("sign", [1, -1])
def test_converter_builds_the_folded_operand_tiles(n, kb, sign):
    rng = np.random.default_rng(n + kb)
    x = rng.uniform(-1, 1, (ROWS, n)).astype(np.float32)

    def box(c0):  # TMA box [c0, c0 + 32) of every row, zero fill outside [0, n)
        out = np.zeros((ROWS, BK), np.float32)
        lo, hi = max(c0, 0), min(c0 + BK, n)
        if hi > lo:
            out[:, lo - c0:hi - c0] = x[:, lo:hi]
        return out

    hi_slot, lo_slot = tma_tile(box(BK * kb)), tma_tile(box(n - BK * (kb + 1)))
    touched = convert_in_place(hi_slot, lo_slot, sign)
    assert np.all(touched == 1)                                   # every chunk of the tile written exactly once
    for i in range(BK):
        m = BK * kb + i
        if m >= n // 2:
            continue                                              # columns the operand B holds zeros for
        want = x[:, m].astype(np.float64) + sign * x[:, n - 1 - m].astype(np.float64)
        j, e = i // 4, i % 4
        got_hi = np.array([hi_slot[r, swz(r, j), e] for r in range(ROWS)])
        got_lo = np.array([lo_slot[r, swz(r, j), e] for r in range(ROWS)])
        assert np.max(np.abs(got_hi.astype(np.float64) + got_lo - want)) <= 2.0 ** -21 * max(1.0, np.max(np.abs(want)))
        assert not np.any(got_hi.view(np.uint32) & 0x1FFF) and not np.any(got_lo.view(np.uint32) & 0x1FFF)


def test_integer_rounding_is_round_to_nearest_ties_away():
    v = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, -(1.0 + 2.0 ** -11), 3.14159274, -2.71828175, 1e-20, 0.0], np.float32)
    got = rna_tf32(v).astype(np.float64)
    step = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(v.astype(np.float64)), 1e-300))) - 10)
    want = np.sign(v) * np.floor(np.abs(v.astype(np.float64)) / step + 0.5) * step   # nearest multiple of the TF32 ulp, ties away
    assert np.array_equal(got[:-1], want[:-1]) and got[-1] == 0.0

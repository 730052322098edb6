"""Host-side model of stft_warp_binmajor_kernel's store phase (zaf-python_b200/csrc/stft.cu): the index arithmetic that
decides which thread writes which (row, frame) of the reference's C-order (N, nt) matrix, replayed in Python.

It pins three properties of the design without a GPU:
  * every element of the matrix is written exactly once, from the right source (direct value or conjugate mirror);
  * a frame is only read from the ring while it is still there (at most 3 frames behind the current tile);
  * away from run boundaries every 16-frame store run starts on a 32-byte sector boundary, whatever nt mod 4 and the
    8-byte phase of the result pointer are -- the property the kernel was built for.
"""
import numpy as np
import pytest

F, SLOTS = 16, 19


def replay(n, nt, phase0, runs_per_clip):
    m = n // 2
    tiles_per_clip = -(-nt // F)
    tiles_per_run = -(-tiles_per_clip // runs_per_clip)
    runs_per_clip = -(-tiles_per_clip // tiles_per_run)
    written = np.zeros((n, nt), np.int32)
    source = np.full((n, nt), -1, np.int64)          # bin whose value (or conjugate) lands here
    conj = np.zeros((n, nt), bool)
    unaligned_interior = 0
    for run in range(runs_per_clip):
        t0 = run * tiles_per_run
        t1 = min(t0 + tiles_per_run, tiles_per_clip)
        jlo, jhi = t0 * F, min(nt, t1 * F)
        ring = {}                                    # slot -> frame
        for t in range(t0, t1 + 1):
            j0 = t * F
            if t < t1:
                for warp in range(F):
                    j = j0 + warp
                    if j < nt:
                        ring[j % SLOTS] = j
            for tid in range(F * 32):
                sw, su = tid & (F - 1), tid // F
                sa = (phase0 + (su & 3) * (nt & 3)) & 3
                sb = (phase0 + ((4 - su) & 3) * (nt & 3)) & 3
                ja = j0 - sa + sw
                if jlo <= ja < jhi:
                    assert ring[ja % SLOTS] == ja, "frame no longer in the ring"
                    for i in range(m // 64):
                        for row, src in ((su + 32 * i, su + 32 * i), (m + su + 32 * i, m + su + 32 * i)):
                            written[row, ja] += 1
                            source[row, ja] = src
                            if sw == 0 and ja > jlo and (phase0 + row * nt + ja) % 4:
                                unaligned_interior += 1
                    if su == 0:
                        for row, c in ((m // 2, False), (m + m // 2, True)):
                            written[row, ja] += 1
                            source[row, ja] = m // 2
                            conj[row, ja] = c
                jb = j0 - sb + sw
                if jlo <= jb < jhi:
                    assert ring[jb % SLOTS] == jb, "frame no longer in the ring"
                    for i in range(m // 64):
                        if i > 0 or su > 0:
                            u = su + 32 * i
                            for row, src in ((n - u, u), (m - u, m + u)):
                                written[row, jb] += 1
                                source[row, jb] = src
                                conj[row, jb] = True
                                if sw == 0 and jb > jlo and (phase0 + row * nt + jb) % 4:
                                    unaligned_interior += 1
    return written, source, conj, unaligned_interior


@pytest.mark.parametrize("n", [256, 512])
@pytest.mark.parametrize("nt", [1, 15, 16, 17, 37, 64, 67])
def test_every_element_once_and_sector_aligned(n, nt):
    rows = np.arange(n)
    for phase0 in range(4):
        for runs in (1, 2, 5):
            written, source, conj, unaligned = replay(n, nt, phase0, runs)
            assert np.all(written == 1), (n, nt, phase0, runs)
            # X[N - k] = conj(X[k]): the source of row r is r itself or its mirror, conjugated exactly when mirrored
            mirror = (n - rows) % n
            for j in range(nt):
                direct = source[:, j] == rows
                assert np.all(direct | (source[:, j] == mirror))
                assert np.all(conj[:, j] == (~direct | (rows == n // 2 + n // 4)))
            assert unaligned == 0, (n, nt, phase0, runs)

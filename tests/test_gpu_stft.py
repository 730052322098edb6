"""GPU parity tests for stft / istft (call through the C ABI; checker = oracle + golden vectors).

Parity metric (north_star "1e-5 relative fp32"):
    max|gpu - ref| <= 1e-5 * max|ref|   and   ||gpu - ref||_2 <= 1e-5 * ||ref||_2
Integer bookkeeping (shapes, frame counts, lengths) must be exactly equal."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def assert_parity(got, ref, tol=TOL):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    mx, l2 = oracle.parity_metrics(got, ref)
    assert mx <= tol and l2 <= tol, (mx, l2)


def test_stft_golden(zaf_gpu, golden):
    g = golden("stft")
    for case in g.cases():
        if not g.has(case, "x"):
            continue
        x, w, hop = g.get(case, "x"), g.get(case, "w"), int(g.get(case, "hop"))
        ref = g.get(case, "stft")
        for layout in ("frame_major", "bin_major"):
            got = zaf_gpu.stft(x.astype(np.float32), w, hop, layout=layout)
            assert got.dtype == np.complex64
            assert_parity(got, ref)
        assert zaf_gpu.stft(x.astype(np.float32), w, hop, layout="bin_major").flags.c_contiguous


def test_istft_golden(zaf_gpu, golden):
    g = golden("stft")
    for case in g.cases():
        w, hop = g.get(case, "w"), int(g.get(case, "hop"))
        spec = g.get(case, "stft") if g.has(case, "stft") else g.get(case, "spec")
        ref = g.get(case, "istft")
        got_c = zaf_gpu.istft(np.ascontiguousarray(spec.astype(np.complex64)), w, hop)           # bin-major memory
        got_f = zaf_gpu.istft(np.ascontiguousarray(spec.T.astype(np.complex64)).T, w, hop)       # frame-major memory
        assert got_c.dtype == np.float32
        assert_parity(got_c, ref)
        assert_parity(got_f, ref)


@pytest.mark.parametrize("n", [4096, 2048, 1024, 512, 256])
@pytest.mark.parametrize("force", [1, 2])
def test_stft_2048_kernels_agree_with_oracle(zaf_gpu, force, n):
    """The warp-per-frame kernel (2; window lengths 512 ... 4096) and the generic Stockham kernel (1) on the same input."""
    rng = np.random.default_rng(20261017 + 2)
    x = rng.uniform(-1, 1, (3, 20000)).astype(np.float32)
    w = oracle.hamming_periodic(n)
    for hop in (n // 4, n // 2, n // 8, n, 300):
        plan, _ = zaf_gpu._stft_plan(w, hop)
        zaf_gpu._lib.check(zaf_gpu._lib.lib().zafb_stft_plan_force_kernel(plan, force))
        try:
            got = zaf_gpu.stft(x, w, hop)
        finally:
            zaf_gpu._lib.lib().zafb_stft_plan_force_kernel(plan, 0)
        for c in range(x.shape[0]):
            assert_parity(got[c], oracle.stft(x[c], w, hop))


@pytest.mark.parametrize("n", [4096, 2048, 1024, 512, 256])
@pytest.mark.parametrize("force", [1, 2])
@pytest.mark.parametrize("ratio", [8, 4, 2])
def test_istft_2048_kernels_agree_with_oracle(zaf_gpu, force, ratio, n):
    """The warp-per-run overlap-add kernel (2) and the generic tile kernel (1) on the same
    NON-Hermitian spectra (the reference keeps Re(ifft) of whatever it is given, zaf.py:223);
    97 frames per clip so that a clip is split into several runs."""
    if n == 256 and ratio == 8 and force == 2:
        pytest.skip("N = 256 has 4 points per lane: the warp kernel serves hop = N/2 and N/4 only")
    hop = n // ratio
    rng = np.random.default_rng(20261017 + hop)
    nt, clips = 97, 3
    spec = (rng.standard_normal((clips, nt, n)) + 1j * rng.standard_normal((clips, nt, n))).astype(np.complex64)
    w = oracle.hamming_periodic(n)
    plan, _ = zaf_gpu._stft_plan(w, hop)
    zaf_gpu._lib.check(zaf_gpu._lib.lib().zafb_stft_plan_force_kernel(plan, force))
    try:
        got = zaf_gpu.istft(np.swapaxes(spec, 1, 2), w, hop)  # frame-major memory, (clips, N, nt) view
        one = zaf_gpu.istft(np.swapaxes(spec, 1, 2)[1], w, hop)
    finally:
        zaf_gpu._lib.lib().zafb_stft_plan_force_kernel(plan, 0)
    assert got.shape == (clips, nt * hop - (n - hop))
    for c in range(clips):
        assert_parity(got[c], oracle.istft(spec[c].T.astype(np.complex128), w, hop))
    assert np.array_equal(one, got[1])  # bitwise: a clip's result does not depend on the batch around it


def test_istft_2048_short_inputs(zaf_gpu):
    """nt < N/hop gives an empty signal, nt == N/hop exactly one hop-block (zaf.py:236-238)."""
    rng = np.random.default_rng(5)
    n, hop = 2048, 512
    w = oracle.hamming_periodic(n)
    for nt in (1, 3, 4, 5, 9):
        spec = (rng.standard_normal((n, nt)) + 1j * rng.standard_normal((n, nt))).astype(np.complex64)
        ref = oracle.istft(spec.astype(np.complex128), w, hop)
        got = zaf_gpu.istft(np.ascontiguousarray(spec.T).T, w, hop)
        assert got.shape == ref.shape
        if ref.size:
            assert_parity(got, ref)


@pytest.mark.parametrize("n,hop,ns", [(2048, 512, 48000), (2048, 1024, 480000), (1024, 256, 80000), (512, 128, 5000),
                                      (256, 64, 1000), (64, 16, 0), (128, 32, 37), (4096, 1024, 30000), (8, 2, 50),
                                      (2, 1, 9), (2048, 300, 9000), (2048, 511, 9000), (1024, 1024, 4000)])
def test_stft_istft_vs_oracle(zaf_gpu, n, hop, ns):
    rng = np.random.default_rng(n * 7 + hop)
    x = rng.uniform(-1, 1, ns).astype(np.float32)
    w = oracle.hamming_periodic(n)
    ref = oracle.stft(x, w, hop)
    got = zaf_gpu.stft(x, w, hop)
    assert_parity(got, ref)
    y_ref = oracle.istft(ref, w, hop)
    y = zaf_gpu.istft(got, w, hop)
    assert_parity(y, y_ref)


@pytest.mark.parametrize("n,hop,ns", [(63, 10, 200), (100, 25, 1000), (255, 64, 3000), (1, 1, 5), (3, 1, 10), (1000, 250, 5000)])
def test_stft_istft_non_power_of_two(zaf_gpu, n, hop, ns):
    """The reference accepts any window length (pocketfft); the GPU path uses a direct DFT there."""
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, ns).astype(np.float32)
    w = np.hanning(n + 2)[1:-1] if n > 1 else np.ones(1)
    ref = oracle.stft(x, w, hop)
    got = zaf_gpu.stft(x, w, hop)
    assert_parity(got, ref)
    assert_parity(zaf_gpu.istft(got, w, hop), oracle.istft(ref, w, hop))


def test_known_answers(zaf_gpu):
    n, hop = 2048, 512
    w = np.ones(n)
    # impulse at the centre of frame 2 -> flat magnitude spectrum in that frame
    x = np.zeros(8192, np.float32)
    x[2 * hop] = 1.0  # frame j covers [j*hop - 1024, j*hop + 1024): sample 2*hop sits at offset 1024 of frame 2
    got = zaf_gpu.stft(x, w, hop)
    assert np.allclose(np.abs(got[:, 2]), 1.0, atol=1e-6)
    # DC
    x = np.ones(8192, np.float32)
    got = zaf_gpu.stft(x, w, hop)
    assert abs(got[0, 4] - n) < 1e-2 and np.max(np.abs(got[1:, 4])) < 1e-2
    # bin-centred cosine: energy only in bins k0 and N-k0
    k0 = 37
    x = np.cos(2 * np.pi * k0 * np.arange(8192) / n).astype(np.float32)
    got = zaf_gpu.stft(x, w, hop)
    col = np.abs(got[:, 5])
    assert abs(col[k0] - n / 2) < 0.05 and abs(col[n - k0] - n / 2) < 0.05
    col[[k0, n - k0]] = 0
    assert col.max() < 0.05
    # Nyquist
    x = np.cos(np.pi * np.arange(8192)).astype(np.float32)
    got = zaf_gpu.stft(x, w, hop)
    assert abs(abs(got[n // 2, 5]) - n) < 1e-2


def test_hermitian_symmetry_and_linearity(zaf_gpu):
    rng = np.random.default_rng(11)
    n, hop = 2048, 512
    w = oracle.hamming_periodic(n)
    a = rng.uniform(-1, 1, (2, 30000)).astype(np.float32)
    sa = zaf_gpu.stft(a, w, hop)
    # X[N-k] == conj(X[k]) up to rounding
    assert np.max(np.abs(sa[:, 1:n // 2, :] - np.conj(sa[:, :n // 2:-1, :]))) <= 2e-6 * np.max(np.abs(sa))
    s_sum = zaf_gpu.stft(a[0] + 2 * a[1], w, hop)
    mx, l2 = oracle.parity_metrics(s_sum, sa[0] + 2 * sa[1])
    assert mx <= 1e-5 and l2 <= 1e-5


def test_round_trip_quirks(zaf_gpu):
    """hop = N/2: identity.  hop = N/4: the reference returns x advanced by N/4 (appendix A.3)."""
    rng = np.random.default_rng(12)
    n = 2048
    w = oracle.hamming_periodic(n)
    x = rng.uniform(-1, 1, 60000).astype(np.float32)
    y = zaf_gpu.istft(zaf_gpu.stft(x, w, n // 2), w, n // 2)
    assert np.max(np.abs(y[:len(x)] - x)) <= 1e-5
    y = zaf_gpu.istft(zaf_gpu.stft(x, w, n // 4), w, n // 4)
    shift = (n - n // 4) - n // 2
    m = min(len(y), len(x) - shift)
    assert np.max(np.abs(y[:m] - x[shift:shift + m])) <= 1e-5


def test_batch_equals_single_and_device_resident(zaf_gpu):
    rng = np.random.default_rng(13)
    n, hop = 2048, 512
    w = oracle.hamming_periodic(n)
    x = rng.uniform(-1, 1, (5, 20000)).astype(np.float32)
    batch = zaf_gpu.stft(x, w, hop)
    assert batch.shape == (5, n, oracle.stft_geometry(20000, n, hop)[1])
    for c in range(5):
        assert np.array_equal(batch[c], zaf_gpu.stft(x[c], w, hop))  # bitwise: no cross-clip arithmetic
    xd = zaf_gpu.to_device(x)
    sd = zaf_gpu.stft(xd, w, hop)
    assert isinstance(sd, zaf_gpu.DeviceArray) and sd.shape == batch.shape
    assert np.array_equal(sd.to_host(), batch)
    yd = zaf_gpu.istft(sd, w, hop)
    assert np.array_equal(yd.to_host(), zaf_gpu.istft(batch, w, hop))


def test_errors(zaf_gpu):
    w = oracle.hamming_periodic(64)
    with pytest.raises(ValueError):
        zaf_gpu.stft(np.zeros((2, 3, 4), np.float32), w, 16)
    with pytest.raises(ValueError):
        zaf_gpu.stft(np.zeros(100, np.float32), w, 16, layout="nope")
    with pytest.raises(ValueError):
        zaf_gpu.istft(np.zeros((32, 5), np.complex64), w, 16)
    with pytest.raises(ValueError):
        zaf_gpu.stft(np.zeros(100, np.float32), w, 0)


def test_full_size_properties(zaf_gpu):
    """BASELINE cfg 2 shape on a slice of the batch (64 clips x 10 s @ 48 kHz): device-resident
    stft -> istft round trip checked through the reference's own shift identity, plus Parseval."""
    rng = np.random.default_rng(20261017 + 2)
    n, hop, ns, clips = 2048, 512, 480000, 64
    w = oracle.hamming_periodic(n)
    x = rng.uniform(-1, 1, (clips, ns)).astype(np.float32)
    xd = zaf_gpu.to_device(x)
    sd = zaf_gpu.stft(xd, w, hop)
    assert sd.shape == (clips, n, 939)
    yd = zaf_gpu.istft(sd, w, hop)
    y = yd.to_host()
    assert y.shape == (clips, 939 * hop - (n - hop))
    shift = (n - hop) - n // 2
    m = min(y.shape[1], ns - shift)
    assert np.max(np.abs(y[:, :m] - x[:, shift:shift + m])) <= 1e-5
    # Parseval on one clip: sum|X|^2 = N * sum (w x)^2 per frame
    spec = sd.to_host()[7]
    ref = oracle.stft(x[7], w, hop)
    assert_parity(spec, ref)


def test_bulk_store_variant(zaf_gpu):
    """ZAFB_STFT_BULK=1 sends the spectrum through shared memory and cp.async.bulk (TMA) stores.  Same formulas as the
    default (direct streaming stores); the compiler contracts the multiply-adds differently, so the two agree to
    rounding, and both meet the parity bar."""
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, (5, 30000)).astype(np.float32)
    w = oracle.hamming_periodic(2048)
    ref = zaf_gpu.stft(x, w, 512).copy()
    os.environ["ZAFB_STFT_BULK"] = "1"
    try:
        got = zaf_gpu.stft(x, w, 512)
    finally:
        del os.environ["ZAFB_STFT_BULK"]
    assert oracle.parity_metrics(got, ref)[0] <= 1e-6
    for c in range(5):
        assert_parity(got[c], oracle.stft(x[c], w, 512))


def test_pinned_result_pool_recycles_blocks(zaf_gpu):
    """Large host results live in page-locked blocks that return to a pool when the array dies: a second call reuses
    the block, a result that is still referenced is never overwritten."""
    rng = np.random.default_rng(12)
    x = rng.uniform(-1, 1, (4, 48000)).astype(np.float32)
    w = oracle.hamming_periodic(2048)
    a = zaf_gpu.stft(x, w, 512)
    keep = a.copy()
    addr_a = a.__array_interface__["data"][0]
    b = zaf_gpu.stft(2 * x, w, 512)          # `a` is alive: must land in a different block
    assert b.__array_interface__["data"][0] != addr_a
    assert np.array_equal(a, keep)
    del a, b
    import gc

    gc.collect()
    c = zaf_gpu.stft(x, w, 512)
    assert np.array_equal(c, keep)
    small = zaf_gpu.stft(x[0, :3000], w, 512)  # below the pool threshold: an ordinary array
    assert_parity(small, oracle.stft(x[0, :3000], w, 512))


def test_pcm16_input_path(zaf_gpu):
    """int16 PCM -> normalised fp32 on the device (zaf.wavread's x / 2**15, zaf.py:1199-1202): exact for every sample,
    planar per channel or the channel mean, and directly usable by the batched transforms."""
    rng = np.random.default_rng(21)
    pcm = rng.integers(-32768, 32768, (30001, 2), dtype=np.int16)
    pcm[0] = (-32768, 32767)
    ref = pcm / pow(2, pcm.itemsize * 8 - 1)              # the reference's arithmetic, float64
    planar = zaf_gpu.from_pcm16(pcm)
    assert planar.shape == (2, 30001)
    host = planar.to_host()
    assert host.dtype == np.float32 and np.array_equal(host.astype(np.float64), ref.T)
    mono = zaf_gpu.from_pcm16(pcm, mono=True).to_host()
    assert np.array_equal(mono.astype(np.float64), np.mean(ref, 1))
    assert np.array_equal(zaf_gpu.from_pcm16(pcm[:, 0]).to_host().astype(np.float64), ref[:, 0])
    w = oracle.hamming_periodic(2048)
    spec = zaf_gpu.stft(planar, w, 512).to_host()
    for c in range(2):
        assert_parity(spec[c], oracle.stft(ref[:, c], w, 512))
    with pytest.raises(ValueError):
        zaf_gpu.from_pcm16(pcm.astype(np.int32))


def test_device_batches_with_odd_length_rows(zaf_gpu):
    """to_device pads the row pitch of an odd-length batch to an even number of samples, so device-resident batches
    run on the same vectorised kernels as the host path (bit-identical results), and to_host drops the padding."""
    rng = np.random.default_rng(31)
    x = rng.uniform(-1, 1, (4, 30011)).astype(np.float32)
    xd = zaf_gpu.to_device(x)
    assert xd.shape == (4, 30011) and xd.pitch == 30012
    assert np.array_equal(xd.to_host(), x)
    w = oracle.hamming_periodic(2048)
    assert np.array_equal(zaf_gpu.stft(xd, w, 512).to_host(), zaf_gpu.stft(x, w, 512))
    wk = oracle.kbd_window(2048)
    md = zaf_gpu.mdct(xd, wk)
    assert np.array_equal(md.to_host(), zaf_gpu.mdct(x, wk))
    back = zaf_gpu.imdct(md, wk)            # odd output length M(nt-1)-1: padded pitch again
    assert np.array_equal(back.to_host(), zaf_gpu.imdct(zaf_gpu.mdct(x, wk), wk))


@pytest.mark.parametrize("n", [2048, 1024, 512, 256])
def test_stft_bin_major_direct_kernel(zaf_gpu, monkeypatch, n):
    """layout="bin_major" on the warp-kernel window lengths is written directly by stft_warp_binmajor_kernel (a CTA walks
    a clip in tiles of 16 frames kept in a shared-memory ring; every row is stored through its own sector-aligned
    window).  It must equal the frame-major result bit for bit (same arithmetic) for frame counts of every residue
    mod 4 and mod 16, single-frame clips, few and many clips (whole-clip runs and split runs), and for result buffers at
    every sector phase; ZAFB_STFT_BM_DIRECT=0 (scratch + transpose route) and streaming stores must give the same bits."""
    rng = np.random.default_rng(20261017 + n)
    w = oracle.hamming_periodic(n)
    cases = [((5, 30010), n // 4), ((2, 9000), n // 2), ((3, 40), n // 8), ((1, 12346), 300), ((7, 2 * n), n // 4),
             ((700, 9 * n // 4), n // 4)]
    cases += [((3, 33 * n // 4 + r * n // 4), n // 4) for r in range(4)]   # nt mod 4 = every residue
    for shape, hop in cases:
        x = rng.uniform(-1, 1, shape).astype(np.float32)
        ref = zaf_gpu.stft(x, w, hop)
        for env in ({}, {"ZAFB_STFT_BM_CS": "1"}, {"ZAFB_STFT_BM_DIRECT": "0"}, {"ZAFB_STFT_BM_RUNS_PER_CLIP": "1"},
                    {"ZAFB_STFT_BM_RUNS_PER_CLIP": "2"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = zaf_gpu.stft(x, w, hop, layout="bin_major")
            for k in env:
                monkeypatch.delenv(k)
            assert got.flags.c_contiguous and got.shape == ref.shape
            assert np.array_equal(got, ref), (shape, hop, env)
        assert_parity(ref[-1], oracle.stft(x[-1], w, hop))
    # device-resident input and output (even clip pitch): the result stays in C order on the device; result pointers
    # at every 8-byte phase of a 32-byte sector
    x = rng.uniform(-1, 1, (4, 20000)).astype(np.float32)
    ref = zaf_gpu.stft(x, w, n // 4)
    xd = zaf_gpu.to_device(x)
    sd = zaf_gpu.stft(xd, w, n // 4, layout="bin_major")
    assert not sd.transposed and np.array_equal(sd.to_host(), ref)
    plan, _ = zaf_gpu._stft_plan(w, n // 4)
    lib = zaf_gpu._lib.lib()
    count = int(np.prod(ref.shape))
    for off in (1, 2, 3):
        buf = zaf_gpu.empty((count + 4,), np.complex64)
        zaf_gpu._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), 4, 20000, xd.pitch, C.c_void_p(buf.ptr + 8 * off), 1, None))
        zaf_gpu.synchronize()
        assert np.array_equal(buf.to_host()[off:off + count].reshape(ref.shape), ref), off


def test_stft_host_half_spectrum_path(zaf_gpu, monkeypatch):
    """Large frame-major host results cross PCIe as bins 0 .. N/2 only; host threads write the Hermitian mirror
    (stft_host_mirrored).  Forced on small inputs here: the result must equal the full-copy pipeline bit for bit, for
    several chunks, ragged last chunks, odd signal lengths, pinned and pageable result memory, 1 .. 5 fill threads."""
    rng = np.random.default_rng(77)
    for n, hop, shape in ((2048, 512, (9, 30011)), (1024, 256, (5, 8000)), (512, 128, (3, 999)), (4096, 1024, (4, 20000)),
                          (64, 16, (7, 1000)), (100, 25, (6, 3000))):
        w = oracle.hamming_periodic(n) if n != 100 else np.hanning(102)[1:-1]
        x = rng.uniform(-1, 1, shape).astype(np.float32)
        monkeypatch.setenv("ZAFB_HOST_MIRROR", "0")
        ref = zaf_gpu.stft(x, w, hop)
        monkeypatch.setenv("ZAFB_HOST_MIRROR", "1")
        monkeypatch.setenv("ZAFB_HOST_MIRROR_MIN_MB", "0")
        monkeypatch.setenv("ZAFB_PIPE_CHUNK_MB", "1")
        for threads in (2, 3, 5):
            monkeypatch.setenv("ZAFB_HOST_MIRROR_THREADS", str(threads))
            for layout in ("frame_major", "bin_major"):
                got = zaf_gpu.stft(x, w, hop, layout=layout)
                assert np.array_equal(np.ascontiguousarray(got).view(np.uint32),
                                      np.ascontiguousarray(ref).view(np.uint32)), (n, hop, shape, threads, layout)
        nt = ref.shape[-1]
        pin = zaf_gpu.PinnedArray((shape[0], nt, n), np.complex64)
        pin.array[:] = 0
        got = zaf_gpu.stft(x, w, hop, out=pin.array)
        assert np.array_equal(np.ascontiguousarray(got), np.ascontiguousarray(ref))
        pin.free()
        for k in ("ZAFB_HOST_MIRROR", "ZAFB_HOST_MIRROR_MIN_MB", "ZAFB_PIPE_CHUNK_MB", "ZAFB_HOST_MIRROR_THREADS"):
            monkeypatch.delenv(k)
        assert_parity(ref[-1], oracle.stft(x[-1], w, hop))


@pytest.mark.parametrize("n", [2048, 1024])
@pytest.mark.parametrize("ratio", [2, 4])
def test_istft_bin_major_direct_kernel(zaf_gpu, monkeypatch, n, ratio):
    """istft of a spectrum in the reference's C-order memory is read directly by istft_binmajor_kernel (16-frame tiles,
    the spectrum combined to X[k] + conj(X[N-k]) while loading), on NON-Hermitian spectra, frame counts around the tile
    size, several clips per CTA, clips split into runs."""
    hop = n // ratio
    rng = np.random.default_rng(n + ratio)
    w = oracle.hamming_periodic(n)
    for clips, nt in ((3, 97), (1, 16), (2, 17), (2, ratio), (1, ratio - 1), (160, 35), (400, 40)):
        spec = (rng.standard_normal((clips, n, nt)) + 1j * rng.standard_normal((clips, n, nt))).astype(np.complex64)  # C order
        direct = zaf_gpu.istft(spec, w, hop)
        monkeypatch.setenv("ZAFB_ISTFT_BM_RUNS_PER_CLIP", "3")   # the same clips cut into three runs with warm-up frames
        assert np.array_equal(zaf_gpu.istft(spec, w, hop), direct)
        monkeypatch.delenv("ZAFB_ISTFT_BM_RUNS_PER_CLIP")
        frame_major = zaf_gpu.istft(np.swapaxes(np.ascontiguousarray(np.swapaxes(spec, 1, 2)), 1, 2), w, hop)
        assert direct.shape == (clips, max(0, nt * hop - (n - hop)))
        # same arithmetic, same summation order; the two kernels differ by the compiler's choice of fused multiply-adds
        # (last-bit differences), so the comparison is at a few fp32 ulps of the signal's peak rather than bitwise
        if direct.size:
            peak = float(np.max(np.abs(frame_major)))
            assert np.max(np.abs(direct - frame_major)) <= 4e-7 * peak, (clips, nt)
        monkeypatch.setenv("ZAFB_ISTFT_BM_DIRECT", "0")
        assert np.array_equal(zaf_gpu.istft(spec, w, hop), frame_major)  # the route through transposed scratch IS bitwise
        monkeypatch.delenv("ZAFB_ISTFT_BM_DIRECT")
        # a clip's result does not depend on the batch around it
        if clips > 1 and direct.size:
            assert np.array_equal(zaf_gpu.istft(spec[clips - 1], w, hop), direct[clips - 1])
        if direct.size:
            assert_parity(direct[clips - 1], oracle.istft(spec[clips - 1].astype(np.complex128), w, hop))
    # device-resident C-order spectrum straight from the direct STFT kernel
    x = rng.uniform(-1, 1, (4, 30000)).astype(np.float32)
    sd = zaf_gpu.stft(zaf_gpu.to_device(x), w, hop, layout="bin_major")
    back = zaf_gpu.istft(sd, w, hop).to_host()
    assert np.max(np.abs(back - zaf_gpu.istft(zaf_gpu.stft(x, w, hop), w, hop))) <= 4e-7

"""GPU parity tests for mdct/imdct, dct/dst, melspectrogram/mfcc, cqtspectrogram/cqtchromagram.
Checker = golden vectors of the unmodified reference + the float64 oracle.  Same metric as test_gpu_stft."""
import numpy as np
import pytest
import scipy.sparse

import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def assert_parity(got, ref, tol=TOL, what=""):
    assert got.shape == ref.shape, (got.shape, ref.shape, what)
    mx, l2 = oracle.parity_metrics(got, ref)
    assert mx <= tol and l2 <= tol, (mx, l2, what)


# ------------------------------------------------------------------------------- MDCT / IMDCT
def test_mdct_imdct_golden(zaf_gpu, golden):
    g = golden("mdct")
    for case in g.cases():
        x, w = g.get(case, "x").astype(np.float32), g.get(case, "w")
        ref = g.get(case, "mdct")
        for layout in ("frame_major", "bin_major"):
            got = zaf_gpu.mdct(x, w, layout=layout)
            assert got.dtype == np.float32
            assert_parity(got, ref)
        y_ref = g.get(case, "imdct")
        assert_parity(zaf_gpu.imdct(np.ascontiguousarray(ref.astype(np.float32)), w), y_ref)
        assert_parity(zaf_gpu.imdct(np.ascontiguousarray(ref.T.astype(np.float32)).T, w), y_ref)


@pytest.mark.parametrize("n,ns", [(2048, 44100), (2048, 1323000), (1024, 3000), (256, 1000), (64, 320), (64, 0), (8, 50),
                                  (4, 9), (2, 7), (100, 1000), (6, 40), (1000, 5000), (4096, 20000)])
def test_mdct_imdct_vs_oracle(zaf_gpu, n, ns):
    rng = np.random.default_rng(n + ns)
    x = rng.uniform(-1, 1, ns).astype(np.float32)
    w = oracle.kbd_window(n) if n >= 64 and n % 4 == 0 else oracle.sine_window(n)
    ref = oracle.mdct(x, w)
    got = zaf_gpu.mdct(x, w)
    assert_parity(got, ref)
    y_ref = oracle.imdct(ref, w)
    y = zaf_gpu.imdct(got, w)
    assert y.shape == y_ref.shape == (max(0, (n // 2) * (ref.shape[1] - 1) - 1),)
    assert_parity(y, y_ref)
    # TDAC: perfect reconstruction of the input (zaf.py:1098-1109)
    m = min(len(y), ns)
    if m:
        assert np.max(np.abs(y[:m] - x[:m])) <= 1e-5


@pytest.mark.parametrize("n", [4096, 2048, 1024, 512])
@pytest.mark.parametrize("force", [1, 2])
def test_mdct_imdct_2048_kernels_agree_with_oracle(zaf_gpu, force, n):
    """The warp-per-frame MDCT / warp-per-run IMDCT kernels (2; window lengths 2048 and 1024) and the generic kernels
    (1) on the same batch: odd and even clip lengths, clips long enough to be split into several runs."""
    rng = np.random.default_rng(20261017 + 4)
    w = oracle.kbd_window(n)
    lib = zaf_gpu._lib.lib()
    plan, _ = zaf_gpu._mdct_plan(w)
    for ns in (100001, 65536, 1000):
        x = rng.uniform(-1, 1, (3, ns)).astype(np.float32)
        zaf_gpu._lib.check(lib.zafb_mdct_plan_force_kernel(plan, force))
        try:
            got = zaf_gpu.mdct(x, w)
            back = zaf_gpu.imdct(got, w)
            xd = zaf_gpu.to_device(x)   # odd ns: the row pitch is padded to an even number of samples
            back_dev = zaf_gpu.imdct(zaf_gpu.mdct(xd, w), w)
        finally:
            lib.zafb_mdct_plan_force_kernel(plan, 0)
        for c in range(3):
            ref = oracle.mdct(x[c], w)
            assert_parity(got[c], ref)
            assert_parity(back[c], oracle.imdct(ref, w))
        assert back.shape[1] == (n // 2) * (got.shape[2] - 1) - 1
        m = min(back.shape[1], ns)
        assert np.max(np.abs(back[:, :m] - x[:, :m])) <= 1e-5  # TDAC
        assert np.array_equal(back_dev.to_host(), back)


def test_mdct_errors_and_batch(zaf_gpu):
    with pytest.raises(ValueError):
        zaf_gpu.mdct(np.zeros(100, np.float32), np.ones(255))
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (4, 30000)).astype(np.float32)
    w = oracle.kbd_window(2048)
    batch = zaf_gpu.mdct(x, w)
    for c in range(4):
        assert np.array_equal(batch[c], zaf_gpu.mdct(x[c], w))
    xd = zaf_gpu.to_device(x)
    md = zaf_gpu.mdct(xd, w)
    yd = zaf_gpu.imdct(md, w)
    y = yd.to_host()
    assert np.max(np.abs(y[:, :30000] - x)) <= 1e-5


# ------------------------------------------------------------------------------- DCT / DST
def test_dct_dst_golden(zaf_gpu, golden):
    g = golden("dctdst")
    for case in g.cases():
        x = g.get(case, "x").astype(np.float32)
        for t in (1, 2, 3, 4):
            assert_parity(zaf_gpu.dct(x, t), g.get(case, f"dct{t}"))
            assert_parity(zaf_gpu.dst(x, t), g.get(case, f"dst{t}"))
    assert zaf_gpu.dct(np.ones(8, np.float32), 5) is None and zaf_gpu.dst(np.ones(8, np.float32), 0) is None


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 16, 31, 64, 100, 256, 1024, 2048, 4096])
def test_dct_dst_vs_oracle_fft_and_direct(zaf_gpu, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    lib = zaf_gpu._lib.lib()
    for kind, fn, ofn in ((0, zaf_gpu.dct, oracle.dct), (1, zaf_gpu.dst, oracle.dst)):
        for t in (1, 2, 3, 4):
            got = fn(x, t)
            for c in range(3):
                assert_parity(got[c], ofn(x[c], t))
            # the table-driven direct kernel must agree as well
            plan = zaf_gpu._dct_plans.get((kind, t, n), kind, t, n)
            zaf_gpu._lib.check(lib.zafb_dct_plan_force_direct(plan, 1))
            try:
                got_d = fn(x, t)
            finally:
                lib.zafb_dct_plan_force_direct(plan, 0)
            for c in range(3):
                assert_parity(got_d[c], ofn(x[c], t))


@pytest.mark.parametrize("n", [16, 100, 257, 1000, 1024, 2048])
def test_dct_dst_tensor_core_matrix_path(zaf_gpu, n):
    """Types I (any N) and non-power-of-two N of every type go through the tcgen05 3xTF32 GEMM against the
    closed-form matrix; batch sizes that are not multiples of the 128-row tile, strided input."""
    rng = np.random.default_rng(1000 + n)
    lib = zaf_gpu._lib.lib()
    for batch in (130, 9):
        x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
        for kind, fn, ofn in ((0, zaf_gpu.dct, oracle.dct), (1, zaf_gpu.dst, oracle.dst)):
            for t in (1, 2, 3, 4):
                fft_path = t >= 2 and n & (n - 1) == 0
                plan = zaf_gpu._dct_plans.get((kind, t, n), kind, t, n)
                if fft_path:  # the matrix path does not exist there and must say so
                    zaf_gpu._lib.check(lib.zafb_dct_plan_force_direct(plan, 2))
                    try:
                        with pytest.raises(NotImplementedError):
                            fn(x, t)
                    finally:
                        lib.zafb_dct_plan_force_direct(plan, 0)
                    continue
                zaf_gpu._lib.check(lib.zafb_dct_plan_force_direct(plan, 2))
                try:
                    got = fn(x, t)
                finally:
                    lib.zafb_dct_plan_force_direct(plan, 0)
                for c in (0, batch // 2, batch - 1):
                    assert_parity(got[c], ofn(x[c], t))
                assert np.array_equal(got, fn(x, t))  # the default route for batch >= 8 is the same path


@pytest.mark.parametrize("n", [16, 40, 257, 1000, 1024])
def test_dct_dst_cta_pair_matrix_path(zaf_gpu, n, monkeypatch):
    """Batches of 256 vectors and more take the CTA-pair form of the matrix path (tcgen05 cta_group::2, 256 x 256 tiles,
    persistent over M blocks): a batch that is not a multiple of the 256-row tile, a batch with more M blocks than
    resident pairs (the persistent loop), every type that has a matrix (even/odd form for types I / II, dense for the
    rest); against the oracle and against the single-CTA kernel."""
    rng = np.random.default_rng(2000 + n)
    lib = zaf_gpu._lib.lib()
    for batch in (700, 74 * 256 * 2 + 300):
        x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
        for kind, fn, ofn in ((0, zaf_gpu.dct, oracle.dct), (1, zaf_gpu.dst, oracle.dst)):
            for t in (1, 2, 3, 4):
                if t >= 2 and n & (n - 1) == 0:
                    continue  # FFT path, no matrix
                if batch > 700 and t >= 3:
                    continue  # the large batch once per operand form is enough
                plan = zaf_gpu._dct_plans.get((kind, t, n), kind, t, n)
                zaf_gpu._lib.check(lib.zafb_dct_plan_force_direct(plan, 2))
                try:
                    got = fn(x, t)
                    monkeypatch.setenv("ZAFB_GEMM_PAIR", "0")
                    single = fn(x, t)
                    monkeypatch.delenv("ZAFB_GEMM_PAIR")
                finally:
                    lib.zafb_dct_plan_force_direct(plan, 0)
                for c in (0, 127, 128, 255, 256, batch // 2, batch - 1):
                    assert_parity(got[c], ofn(x[c], t), what=f"kind {kind} type {t} batch {batch} vec {c}")
                # same chains, same chunk sums: the two kernels agree to the last bits over the whole batch
                assert np.max(np.abs(got - single)) <= 2e-6 * np.max(np.abs(single)), (kind, t, batch)


def test_dct_dst_1024_warp_kernel(zaf_gpu):
    """N = 1024, types II-IV: the one-warp-per-vector kernel (forced), the block FFT kernel (forced) and the default
    route against the oracle; a batch that does not fill the last CTA; bit-identical results between calls."""
    rng = np.random.default_rng(77)
    lib = zaf_gpu._lib.lib()
    batch = 2 * 148 * 8 + 13
    x = rng.uniform(-1, 1, (batch, 1024)).astype(np.float32)
    x[1] = 0.0
    x[2, :] = 1.0
    for kind, fn, ofn in ((0, zaf_gpu.dct, oracle.dct), (1, zaf_gpu.dst, oracle.dst)):
        for t in (2, 3, 4):
            plan = zaf_gpu._dct_plans.get((kind, t, 1024), kind, t, 1024)
            res = {}
            for force in (4, 3, 0):
                zaf_gpu._lib.check(lib.zafb_dct_plan_force_direct(plan, force))
                try:
                    res[force] = fn(x, t)
                finally:
                    lib.zafb_dct_plan_force_direct(plan, 0)
                for c in (0, 1, 2, 3, batch // 2, batch - 1):
                    assert_parity(res[force][c], ofn(x[c], t), what=f"kind {kind} type {t} force {force} vec {c}")
            assert np.array_equal(res[4], res[0])  # the warp kernel is the default route
            assert np.array_equal(fn(x[5], t), res[0][5])  # a single vector goes the same way


def test_dst_inverse_pairs(zaf_gpu):
    x = np.random.default_rng(5).uniform(-1, 1, 1024).astype(np.float32)
    assert np.max(np.abs(zaf_gpu.dst(zaf_gpu.dst(x, 1), 1) - x)) <= 1e-5
    assert np.max(np.abs(zaf_gpu.dst(zaf_gpu.dst(x, 2), 3) - x)) <= 1e-5
    assert np.max(np.abs(zaf_gpu.dst(zaf_gpu.dst(x, 4), 4) - x)) <= 1e-5
    assert np.max(np.abs(zaf_gpu.dct(zaf_gpu.dct(x, 2), 3) - x)) <= 1e-5
    assert np.max(np.abs(zaf_gpu.dct(zaf_gpu.dct(x, 1), 1) - x)) <= 1e-5


# ------------------------------------------------------------------------------- mel / mfcc
def test_mel_mfcc_golden(zaf_gpu, golden):
    g = golden("mel")
    for case in g.cases():
        x, w, hop = g.get(case, "x").astype(np.float32), g.get(case, "w"), int(g.get(case, "hop"))
        fb = zaf_gpu.melfilterbank(int(g.get(case, "fs")), len(w), int(g.get(case, "nmel")))
        ncoef = int(g.get(case, "ncoef"))
        ref_mel, ref_mfcc = g.get(case, "mel"), g.get(case, "mfcc")
        for layout in ("frame_major", "bin_major"):
            got = zaf_gpu.melspectrogram(x, w, hop, fb, layout=layout)
            assert got.shape == ref_mel.shape
            if case == "silent":
                assert np.max(np.abs(got)) == 0.0
            else:
                assert_parity(got, ref_mel)
            got = zaf_gpu.mfcc(x, w, hop, fb, ncoef, layout=layout)
            assert got.shape == ref_mfcc.shape
            if case == "silent":
                assert np.max(np.abs(got)) <= 1e-4  # log(eps) is constant -> every kept coefficient is ~0
            else:
                assert_parity(got, ref_mfcc)


def test_mel_mfcc_cfg3_batch_vs_oracle(zaf_gpu):
    """BASELINE cfg 3 shape (5 s @ 16 kHz, N=1024, hop=256, 128 mel, 40 coefficients), a few clips."""
    rng = np.random.default_rng(20261017 + 3)
    x = rng.uniform(-1, 1, (6, 80000)).astype(np.float32)
    w = oracle.hamming_periodic(1024)
    fb = zaf_gpu.melfilterbank(16000, 1024, 128)
    mel = zaf_gpu.melspectrogram(x, w, 256, fb)
    cep = zaf_gpu.mfcc(x, w, 256, fb, 40)
    assert mel.shape == (6, 128, 314) and cep.shape == (6, 40, 314)
    dense = fb.toarray()
    for c in (0, 5):
        assert_parity(mel[c], oracle.melspectrogram(x[c], w, 256, dense))
        assert_parity(cep[c], oracle.mfcc(x[c], w, 256, dense, 40))
    # dense (non-banded) operator and more coefficients than mel rows - 1
    dense_fb = np.abs(rng.standard_normal((10, 512)))
    assert_parity(zaf_gpu.melspectrogram(x[0], w, 256, dense_fb), oracle.melspectrogram(x[0], w, 256, dense_fb))
    got = zaf_gpu.mfcc(x[0], w, 256, dense_fb, 40)
    ref = oracle.mfcc(x[0], w, 256, dense_fb, 40)
    assert got.shape == ref.shape == (9, 314)
    assert_parity(got, ref)


@pytest.mark.parametrize("n_mels,ncoef,fs", [(128, 40, 16000), (40, 13, 16000), (77, 60, 44100), (5, 3, 16000)])
def test_mel_mfcc_tensor_core_route(zaf_gpu, n_mels, ncoef, fs):
    """route="tensor": the filterbank (and the MFCC DCT) as dense 3xTF32 products on the tcgen05 tensor cores.
    Same parity bar as the fused route; frame counts that are not multiples of the 128-row tile; more clips than
    one 16 384-frame chunk holds."""
    rng = np.random.default_rng(4242 + n_mels)
    w = oracle.hamming_periodic(1024)
    fb = zaf_gpu.melfilterbank(fs, 1024, n_mels)
    dense = fb.toarray()
    for clips, ns, hop in ((3, 20001, 256), (70, 80000, 256), (2, 5000, 512)):
        x = rng.uniform(-1, 1, (clips, ns)).astype(np.float32)
        x[0, 1000:3000] = 0.0
        mel = zaf_gpu.melspectrogram(x, w, hop, fb, route="tensor")
        cep = zaf_gpu.mfcc(x, w, hop, fb, ncoef, route="tensor")
        for c in sorted({0, clips // 2, clips - 1}):
            assert_parity(mel[c], oracle.melspectrogram(x[c], w, hop, dense), what=f"mel clip {c}")
            assert_parity(cep[c], oracle.mfcc(x[c], w, hop, dense, ncoef), what=f"mfcc clip {c}")
        fused = zaf_gpu.melspectrogram(x, w, hop, fb)
        assert oracle.parity_metrics(mel, fused)[0] <= 2e-6
    with pytest.raises(NotImplementedError):  # the dense route exists for window_length 1024 only
        zaf_gpu.melspectrogram(x[0], oracle.hamming_periodic(2048), 512, zaf_gpu.melfilterbank(fs, 2048, n_mels), route="tensor")
    with pytest.raises(ValueError):
        zaf_gpu.melspectrogram(x[0], w, 256, fb, route="nope")


@pytest.mark.parametrize("n", [1024, 2048, 512])
@pytest.mark.parametrize("force", [1, 2])
@pytest.mark.parametrize("n_mels,ncoef,fs", [(128, 40, 16000), (40, 13, 16000), (77, 60, 44100), (100, 99, 22050), (1, 1, 16000),
                                             (128, 20, 44100)])
def test_mel_mfcc_1024_kernels_agree_with_oracle(zaf_gpu, force, n_mels, ncoef, fs, n):
    """Warp-per-frame kernel (2) and generic kernel (1) for N = 1024: row counts that are not
    multiples of 32, odd row counts (middle row of the DCT symmetry), more coefficients than 32."""
    rng = np.random.default_rng(n_mels * 1000 + ncoef)
    x = rng.uniform(-1, 1, (3, 20001)).astype(np.float32)
    x[2, 5000:9000] = 0.0  # a stretch of digital silence inside a clip
    w = oracle.hamming_periodic(n)
    fb = zaf_gpu.melfilterbank(fs, n, n_mels)
    dense = fb.toarray()
    lib = zaf_gpu._lib.lib()
    for hop in (n // 4, n // 2):
        plans = [zaf_gpu._mel_plan(w, hop, fb, 0)[0], zaf_gpu._mel_plan(w, hop, fb, ncoef)[0]]
        for p in plans:
            zaf_gpu._lib.check(lib.zafb_mel_plan_force_kernel(p, force))
        try:
            mel = zaf_gpu.melspectrogram(x, w, hop, fb)
            if force == 2 and min(ncoef, n_mels - 1) > 64:  # outside the warp kernel's envelope: must refuse, not guess
                with pytest.raises(NotImplementedError):
                    zaf_gpu.mfcc(x, w, hop, fb, ncoef)
                lib.zafb_mel_plan_force_kernel(plans[1], 0)
            cep = zaf_gpu.mfcc(x, w, hop, fb, ncoef)
        finally:
            for p in plans:
                lib.zafb_mel_plan_force_kernel(p, 0)
        for c in range(3):
            assert_parity(mel[c], oracle.melspectrogram(x[c], w, hop, dense))
            ref = oracle.mfcc(x[c], w, hop, dense, ncoef)
            assert cep[c].shape == ref.shape
            if ref.size:
                assert_parity(cep[c], ref)


# ------------------------------------------------------------------------------- CQT
def _kernel(g, tag):
    shape = tuple(int(s) for s in g.get(tag, "kernel_shape"))
    return scipy.sparse.csr_matrix(
        (g.get(tag, "kernel_data"), g.get(tag, "kernel_indices"), g.get(tag, "kernel_indptr")), shape=shape)


def test_cqt_golden(zaf_gpu, golden):
    g = golden("cqt")
    for tag in g.cases():
        fs, res, _, _ = g.get(tag, "params")
        k = _kernel(g, tag)
        x = g.get(tag, "x").astype(np.float32)
        tr = int(g.get(tag, "time_resolution"))
        for layout in ("frame_major", "bin_major"):
            assert_parity(zaf_gpu.cqtspectrogram(x, int(fs), tr, k, layout=layout), g.get(tag, "spec"))
            assert_parity(zaf_gpu.cqtchromagram(x, int(fs), tr, int(res), k, layout=layout), g.get(tag, "chroma"))


@pytest.mark.parametrize("force", [1, 2, 3])
def test_cqt_32768_kernels_agree_with_oracle(zaf_gpu, force):
    """Even/odd kernel (3, the default), register-FFT kernel (2) and generic kernel (1) at fft_length 32768 (BASELINE cfg 5 kernel): odd clip
    length (unaligned rows), edge frames that need zero padding, chroma fold, a mirrored band and the Nyquist column."""
    rng = np.random.default_rng(20261017 + 5)
    fs = 44100
    k = zaf_gpu.cqtkernel(fs, 12, 32.70319566257483, 4186.009044809578)
    assert k.shape == (84, 32768)
    x = rng.uniform(-1, 1, (2, 30001)).astype(np.float32)
    lib = zaf_gpu._lib.lib()
    step = round(fs / 25)
    plan = zaf_gpu._cqt_plan(k, step)[0]
    zaf_gpu._lib.check(lib.zafb_cqt_plan_force_kernel(plan, force))
    try:
        spec = zaf_gpu.cqtspectrogram(x, fs, 25, k)
        chroma = zaf_gpu.cqtchromagram(x, fs, 25, 12, k)
        spec_c = zaf_gpu.cqtspectrogram(x[0], fs, 25, k, layout="bin_major")
    finally:
        lib.zafb_cqt_plan_force_kernel(plan, 0)
    for c in range(2):
        ref = oracle.cqtspectrogram(x[c], fs, 25, k)
        assert_parity(spec[c], ref)
        assert_parity(chroma[c], oracle.cqtchromagram(x[c], fs, 25, 12, k))
    assert np.array_equal(spec_c, spec[0]) and spec_c.flags.c_contiguous
    # arbitrary complex operator: DC column, Nyquist column, a band across L/2 and one in the mirrored half
    L = 32768
    dense = scipy.sparse.lil_matrix((4, L), dtype=complex)
    dense[0, 0:40] = rng.standard_normal(40) + 1j * rng.standard_normal(40)
    dense[1, L // 2 - 30:L // 2 + 30] = rng.standard_normal(60) + 1j * rng.standard_normal(60)
    dense[2, L - 500:L - 400] = rng.standard_normal(100)
    dense[3, 8000:8400] = 1j * rng.standard_normal(400)
    kk = scipy.sparse.csr_matrix(dense)
    plan = zaf_gpu._cqt_plan(kk, 4410)[0]
    if force == 3:  # a complex kernel with mirrored bands is outside the even/odd kernel's contract
        with pytest.raises(NotImplementedError):
            zaf_gpu._lib.check(lib.zafb_cqt_plan_force_kernel(plan, force))
            zaf_gpu.cqtspectrogram(x[1], fs, 10, kk)
        lib.zafb_cqt_plan_force_kernel(plan, 0)
        # a real kernel that uses DC, odd / even band edges and the last column below L/2
        real = scipy.sparse.lil_matrix((5, L), dtype=complex)
        real[0, 0:41] = rng.standard_normal(41)
        real[1, 7:8] = 1.5
        real[2, 4001:4400] = rng.standard_normal(399)
        real[3, L // 2 - 257:L // 2] = rng.standard_normal(257)
        real[4, 100:2000] = rng.standard_normal(1900)
        kk = scipy.sparse.csr_matrix(real)
        plan = zaf_gpu._cqt_plan(kk, 4410)[0]
    zaf_gpu._lib.check(lib.zafb_cqt_plan_force_kernel(plan, force))
    try:
        got = zaf_gpu.cqtspectrogram(x[1], fs, 10, kk)
    finally:
        lib.zafb_cqt_plan_force_kernel(plan, 0)
    assert_parity(got, oracle.cqtspectrogram(x[1], fs, 10, kk))


def test_cqt_tensor_core_route(zaf_gpu):
    """route="tensor": the CQT kernel as a dense 3xTF32 product on the tcgen05 tensor cores (BASELINE cfg 5: "sparse CQT
    kernel as packed tensor-core GEMM").  The cfg 5 kernel (84 rows, one N tile) and the reference's example kernel (144
    rows, two N tiles); spectrogram and chromagram; both layouts; same parity bar as the fused route."""
    rng = np.random.default_rng(20261017 + 55)
    fs = 44100
    x = rng.uniform(-1, 1, (3, 40001)).astype(np.float32)
    for res, fmin, fmax, rows in ((12, 32.70319566257483, 4186.009044809578, 84), (24, 55.0, 3520.0, 144)):
        k = zaf_gpu.cqtkernel(fs, res, fmin, fmax)
        assert k.shape == (rows, 32768)
        spec = zaf_gpu.cqtspectrogram(x, fs, 25, k, route="tensor")
        chroma = zaf_gpu.cqtchromagram(x, fs, 25, res, k, route="tensor")
        fused = zaf_gpu.cqtspectrogram(x, fs, 25, k)
        assert oracle.parity_metrics(spec, fused)[0] <= 2e-6
        for c in range(3):
            assert_parity(spec[c], oracle.cqtspectrogram(x[c], fs, 25, k), what=f"spec {rows} {c}")
            assert_parity(chroma[c], oracle.cqtchromagram(x[c], fs, 25, res, k), what=f"chroma {rows} {c}")
        spec_c = zaf_gpu.cqtspectrogram(x[1], fs, 25, k, layout="bin_major", route="tensor")
        assert spec_c.flags.c_contiguous and np.array_equal(spec_c, spec[1])
    # a complex kernel, or one with a band in the mirrored half, has no dense real operand: the route must refuse
    dense = scipy.sparse.lil_matrix((2, 32768), dtype=complex)
    dense[0, 100:200] = 1j * rng.standard_normal(100)
    dense[1, 300:400] = rng.standard_normal(100)
    with pytest.raises(NotImplementedError):
        zaf_gpu.cqtspectrogram(x[0], fs, 10, scipy.sparse.csr_matrix(dense), route="tensor")


def test_cqt_small_kernels_vs_oracle(zaf_gpu):
    """Smaller FFT lengths (odd and even log2), columns in the upper half of the spectrum, complex weights."""
    rng = np.random.default_rng(9)
    for fs, res, fmin, fmax, tr in ((8000, 6, 220.0, 1760.0, 50), (4000, 4, 200.0, 1600.0, 40), (16000, 12, 440.0, 3520.0, 25)):
        k = zaf_gpu.cqtkernel(fs, res, fmin, fmax)
        x = rng.uniform(-1, 1, (2, fs)).astype(np.float32)
        got = zaf_gpu.cqtspectrogram(x, fs, tr, k)
        for c in range(2):
            assert_parity(got[c], oracle.cqtspectrogram(x[c], fs, tr, k))
    # an arbitrary complex banded operator, bands crossing L/2
    L, nf = 1024, 5
    dense = np.zeros((nf, L), complex)
    for r in range(nf):
        lo = 400 + 40 * r
        dense[r, lo:lo + 150] = rng.standard_normal(150) + 1j * rng.standard_normal(150)
    dense[0, 0] = 1.0
    x = rng.uniform(-1, 1, 6000).astype(np.float32)
    k = scipy.sparse.csr_matrix(dense)
    assert_parity(zaf_gpu.cqtspectrogram(x, 8000, 100, k), oracle.cqtspectrogram(x, 8000, 100, k))


# ------------------------------------------------------------------------------- layouts
def test_bin_major_layout_on_the_warp_kernels(zaf_gpu, monkeypatch):
    """layout="bin_major" (the reference's C-order memory) on the sizes served by the warp kernels goes through a
    tiled transpose: results must equal the frame-major ones bit for bit, in C-contiguous memory, also when the batch
    is cut into several scratch chunks, and the inverse transforms must accept C-order input."""
    monkeypatch.setenv("ZAFB_TRANSPOSE_CHUNK_MB", "1")
    rng = np.random.default_rng(99)
    x = rng.uniform(-1, 1, (5, 30011)).astype(np.float32)
    w = oracle.hamming_periodic(2048)
    a = zaf_gpu.stft(x, w, 512)
    b = zaf_gpu.stft(x, w, 512, layout="bin_major")
    assert b.flags.c_contiguous and not a.flags.c_contiguous and np.array_equal(a, b)
    # C-order input goes through istft_binmajor_kernel: the same arithmetic in the same order, but a different kernel
    # (the compiler fuses multiply-adds differently): equal to a few fp32 ulps of the peak, not bit for bit
    assert np.max(np.abs(zaf_gpu.istft(b, w, 512) - zaf_gpu.istft(a, w, 512))) <= 4e-7
    assert_parity(b[3], oracle.stft(x[3], w, 512))
    wk = oracle.kbd_window(2048)
    ma = zaf_gpu.mdct(x, wk)
    mb = zaf_gpu.mdct(x, wk, layout="bin_major")
    # C-order MDCT comes from mdct_binmajor_kernel: the arithmetic of the frame-major kernel in another kernel (the
    # compiler fuses multiply-adds differently): equal to an fp32 ulp of the peak, not bit for bit
    assert mb.flags.c_contiguous and np.max(np.abs(ma - mb)) <= 4e-7 * np.max(np.abs(ma))
    assert np.max(np.abs(zaf_gpu.imdct(mb, wk) - zaf_gpu.imdct(ma, wk))) <= 1e-6
    monkeypatch.setenv("ZAFB_MDCT_BM_DIRECT", "0")  # the scratch + transpose route is bit-identical to frame-major
    assert np.array_equal(zaf_gpu.mdct(x, wk, layout="bin_major"), ma)
    monkeypatch.delenv("ZAFB_MDCT_BM_DIRECT")
    w1 = oracle.hamming_periodic(1024)
    fb = zaf_gpu.melfilterbank(16000, 1024, 128)
    for route in ("fused", "tensor"):
        assert np.array_equal(zaf_gpu.melspectrogram(x, w1, 256, fb, layout="bin_major", route=route),
                              zaf_gpu.melspectrogram(x, w1, 256, fb, route=route))
        assert np.array_equal(zaf_gpu.mfcc(x, w1, 256, fb, 40, layout="bin_major", route=route),
                              zaf_gpu.mfcc(x, w1, 256, fb, 40, route=route))
    # device-resident, one chunk (even clip pitch, so the same warp kernels serve it)
    monkeypatch.delenv("ZAFB_TRANSPOSE_CHUNK_MB")
    x2 = np.ascontiguousarray(x[:, :30010])
    a2 = zaf_gpu.stft(x2, w, 512)
    sd = zaf_gpu.stft(zaf_gpu.to_device(x2), w, 512, layout="bin_major")
    assert not sd.transposed and np.array_equal(sd.to_host(), a2)
    assert np.max(np.abs(zaf_gpu.istft(sd, w, 512).to_host() - zaf_gpu.istft(a2, w, 512))) <= 4e-7


@pytest.mark.parametrize("n", [2048, 1024])
def test_mdct_imdct_bin_major_direct_kernels(zaf_gpu, monkeypatch, n):
    """layout="bin_major" MDCT at window lengths 2048 / 1024 is written by mdct_binmajor_kernel (tiles of 32 frames in a
    shared-memory ring, every row stored through its own sector-aligned window) and C-order IMDCT input is read by
    imdct_binmajor_kernel: same arithmetic as the frame-major warp kernels.  Frame counts of every residue mod 8 and
    around the 32-frame tile, single-frame and two-frame clips, whole-clip and split runs, result pointers at every
    4-byte phase of a 32-byte sector; ZAFB_*_BM_DIRECT=0 (scratch + transpose route) must agree.  Another kernel means
    other multiply-add fusions: agreement to a few fp32 ulps (4e-7 of the peak for one transform, 1e-6 for two in a
    row), not bit for bit; the oracle check at 1e-5 is separate."""
    import ctypes as C
    rng = np.random.default_rng(20261017 + n)
    w = oracle.kbd_window(n)
    m = n // 2
    monkeypatch.setenv("ZAFB_IMDCT_BM_MIN_CLIPS", "1")
    lens = [m * k + 2 for k in range(1, 10)] + [m * 31, m * 32 + 10, m * 33 + 4, m * 70 + 6, 10, 2]
    for i, ns in enumerate(lens):
        clips = 3 if i % 2 else 2
        x = rng.uniform(-1, 1, (clips, ns)).astype(np.float32)
        ref = zaf_gpu.mdct(x, w)
        y_ref = zaf_gpu.imdct(ref, w)
        peak = max(1.0, float(np.max(np.abs(ref))))
        for env in ({}, {"ZAFB_MDCT_BM_DIRECT": "0", "ZAFB_IMDCT_BM_DIRECT": "0"}, {"ZAFB_MDCT_BM_RUNS_PER_CLIP": "1"},
                    {"ZAFB_MDCT_BM_RUNS_PER_CLIP": "2"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = zaf_gpu.mdct(x, w, layout="bin_major")
            y = zaf_gpu.imdct(got, w)
            for k in env:
                monkeypatch.delenv(k)
            assert got.flags.c_contiguous and got.shape == ref.shape
            assert np.max(np.abs(got - ref), initial=0.0) <= 4e-7 * peak, (ns, env)
            assert y.shape == y_ref.shape and np.max(np.abs(y - y_ref), initial=0.0) <= 1e-6, (ns, env)
        assert_parity(ref[-1], oracle.mdct(x[-1], w))
        assert_parity(y_ref[-1], oracle.imdct(oracle.mdct(x[-1], w), w))
    # many short clips (every CTA walks several clips), device-resident, result pointers at every sector phase
    x = rng.uniform(-1, 1, (300, 5 * m + 2)).astype(np.float32)
    ref = zaf_gpu.mdct(x, w)
    xd = zaf_gpu.to_device(x)
    sd = zaf_gpu.mdct(xd, w, layout="bin_major")
    assert not sd.transposed
    peak = float(np.max(np.abs(ref)))
    assert np.max(np.abs(sd.to_host() - ref)) <= 4e-7 * peak
    assert np.max(np.abs(zaf_gpu.imdct(sd, w).to_host() - zaf_gpu.imdct(ref, w))) <= 1e-6
    plan, _ = zaf_gpu._mdct_plan(w)
    lib = zaf_gpu._lib.lib()
    count = int(np.prod(ref.shape))
    nt = ref.shape[-1]
    length = zaf_gpu.imdct_geometry(m, nt)[1]
    for off in range(1, 8):
        buf = zaf_gpu.empty((count + 8,), np.float32)
        zaf_gpu._lib.check(lib.zafb_mdct_f32(plan, C.c_void_p(xd.ptr), 300, x.shape[1], xd.pitch, C.c_void_p(buf.ptr + 4 * off), 1, None))
        yb = zaf_gpu.empty((300, length + 1), np.float32)
        zaf_gpu._lib.check(lib.zafb_imdct_f32(plan, C.c_void_p(buf.ptr + 4 * off), 300, nt, 1, C.c_void_p(yb.ptr), length + 1, None))
        zaf_gpu.synchronize()
        assert np.max(np.abs(buf.to_host()[off:off + count].reshape(ref.shape) - ref)) <= 4e-7 * peak, off
        assert np.max(np.abs(yb.to_host()[:, :length] - zaf_gpu.imdct(ref, w))) <= 1e-6, off


def test_sum_of_sinusoids_parity_and_the_fp32_floor_of_mfcc(zaf_gpu):
    """A sum of sinusoids (SURVEY.md section 8d asks for one): every output keeps the 1e-5 bar.  MFCC takes the LOG of
    mel energies that sit 100+ dB below the spectral peak; an fp32 FFT resolves a bin only to about 1e-8 of the peak
    amplitude, so in pure fp32 those energies carry relative errors of 1e-4..1e-3 and the coefficients miss 1e-5
    (precision="float32": measured 2.3e-5).  The default (precision="auto") detects such frames in the fp32 kernel and
    recomputes them in float64 on the GPU: 1e-5 holds without the caller asking for anything."""
    t = np.arange(80000) / 16000.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.1 * np.sin(2 * np.pi * 3000 * t)).astype(np.float32)
    w = oracle.hamming_periodic(1024)
    fb = zaf_gpu.melfilterbank(16000, 1024, 128)
    dense = fb.toarray()
    assert_parity(zaf_gpu.stft(x, w, 256), oracle.stft(x, w, 256))
    assert_parity(zaf_gpu.melspectrogram(x, w, 256, fb), oracle.melspectrogram(x, w, 256, dense))
    wk = oracle.kbd_window(2048)
    assert_parity(zaf_gpu.mdct(x, wk), oracle.mdct(x, wk))
    for t_ in (2, 4):
        assert_parity(zaf_gpu.dct(x[:1024], t_), oracle.dct(x[:1024], t_))
    ref = oracle.mfcc(x, w, 256, dense, 40)
    assert_parity(zaf_gpu.mfcc(x, w, 256, fb, 40), ref)                            # default: automatic float64 re-computation
    assert_parity(zaf_gpu.mfcc(x, w, 256, fb, 40, layout="bin_major"), ref)
    assert_parity(zaf_gpu.mfcc(x, w, 256, fb, 40, precision="float32"), ref, tol=1e-4)  # the fp32 floor, for the record
    assert_parity(zaf_gpu.mfcc(x, w, 256, fb, 40, precision="float64"), ref)
    assert_parity(zaf_gpu.melspectrogram(x, w, 256, fb, precision="float64"), oracle.melspectrogram(x, w, 256, dense))
    # a batch that mixes tonal, broadband and silent clips: only the tonal frames take the float64 path, every clip meets 1e-5
    rng = np.random.default_rng(8)
    batch = np.stack([x[:30000], rng.uniform(-1, 1, 30000).astype(np.float32), np.zeros(30000, np.float32),
                      (x[:30000] * np.float32(1e-3))])
    got = zaf_gpu.mfcc(batch, w, 256, fb, 40)
    for c in range(4):
        assert_parity(got[c], oracle.mfcc(batch[c], w, 256, dense, 40))
    # the broadband clip is bit-identical to the pure fp32 result (its frames never left the fp32 kernel)
    assert np.array_equal(got[1], zaf_gpu.mfcc(batch[1], w, 256, fb, 40, precision="float32"))
    # other window lengths: warp kernels at 2048 / 512 and the generic kernel at 256, float64 re-computation by the block kernel
    for n, hop, fs, mels, nc in ((2048, 1024, 44100, 128, 20), (512, 128, 16000, 64, 13), (256, 64, 8000, 40, 13)):
        tt = np.arange(40000) / fs
        xt = (0.5 * np.sin(2 * np.pi * 440 * tt) + 0.1 * np.sin(2 * np.pi * 1500 * tt)).astype(np.float32)
        wn = oracle.hamming_periodic(n)
        fbn = zaf_gpu.melfilterbank(fs, n, mels)
        assert_parity(zaf_gpu.mfcc(xt, wn, hop, fbn, nc), oracle.mfcc(xt, wn, hop, fbn.toarray(), nc))


@pytest.mark.parametrize("n,hop,n_mels,ncoef,fs", [(1024, 256, 128, 40, 16000), (2048, 1024, 128, 20, 44100), (256, 64, 40, 13, 8000),
                                                  (4096, 1000, 77, 76, 44100)])
def test_mel_mfcc_float64_route(zaf_gpu, n, hop, n_mels, ncoef, fs):
    """precision="float64": any power-of-two window length, both layouts, batches, silent stretches; results within
    2e-7 of the float64 oracle (the only fp32 roundings left are the input, the filterbank weights and the output)."""
    rng = np.random.default_rng(n + n_mels)
    x = rng.uniform(-1, 1, (3, 12001)).astype(np.float32)
    x[1, 2000:6000] = 0.0
    w = oracle.hamming_periodic(n)
    fb = zaf_gpu.melfilterbank(fs, n, n_mels)
    dense = fb.toarray()
    mel = zaf_gpu.melspectrogram(x, w, hop, fb, precision="float64")
    cep = zaf_gpu.mfcc(x, w, hop, fb, ncoef, precision="float64")
    for c in range(3):
        assert_parity(mel[c], oracle.melspectrogram(x[c], w, hop, dense), tol=2e-7)
        ref = oracle.mfcc(x[c], w, hop, dense, ncoef)
        assert cep[c].shape == ref.shape
        assert_parity(cep[c], ref, tol=1e-6)
    assert np.array_equal(zaf_gpu.mfcc(x, w, hop, fb, ncoef, precision="float64", layout="bin_major"), cep)
    with pytest.raises(ValueError):
        zaf_gpu.mfcc(x, w, hop, fb, ncoef, precision="float16")


@pytest.mark.parametrize("n,hop,fs,n_mels,ncoef", [(1000, 250, 16000, 64, 20), (1764, 441, 44100, 128, 40), (300, 75, 8000, 30, 12)])
def test_mel_mfcc_window_lengths_that_are_not_powers_of_two(zaf_gpu, n, hop, fs, n_mels, ncoef):
    """The reference accepts any window length (zaf.py:369 -> stft with pocketfft); here the any-length STFT kernels feed
    mel_from_spectrum_kernel.  Both layouts, batches, host and device inputs."""
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (3, 9001)).astype(np.float32)
    w = np.hanning(n + 2)[1:-1]
    fb = zaf_gpu.melfilterbank(fs, n, n_mels)
    dense = fb.toarray()
    assert dense.shape == (n_mels, n // 2)
    mel = zaf_gpu.melspectrogram(x, w, hop, fb)
    cep = zaf_gpu.mfcc(x, w, hop, fb, ncoef)
    for c in range(3):
        assert_parity(mel[c], oracle.melspectrogram(x[c], w, hop, dense))
        assert_parity(cep[c], oracle.mfcc(x[c], w, hop, dense, ncoef))
    assert np.array_equal(zaf_gpu.melspectrogram(x, w, hop, fb, layout="bin_major"), mel)
    assert np.array_equal(zaf_gpu.mfcc(zaf_gpu.to_device(x), w, hop, fb, ncoef).to_host(), cep)
    with pytest.raises(NotImplementedError):
        zaf_gpu.mfcc(x, w, hop, fb, ncoef, precision="float64")


@pytest.mark.parametrize("n", [512, 2048, 4096])
def test_dct_dst_warp_kernels_512_2048(zaf_gpu, n):
    """The one-warp-per-vector kernel (types II-IV and their DST twins) at N = 512, 2048 and 4096 (r02; N = 1024 has its own
    test): forced through the plan hook, against the oracle, plus the orthonormal inverse pairs."""
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (37, n)).astype(np.float32)
    lib = zaf_gpu._lib.lib()
    for kind, fn, ofn in ((0, zaf_gpu.dct, oracle.dct), (1, zaf_gpu.dst, oracle.dst)):
        for t in (2, 3, 4):
            plan = zaf_gpu._dct_plans.get((kind, t, n), kind, t, n)
            zaf_gpu._lib.check(lib.zafb_dct_plan_force_direct(plan, 4))  # 4 = require the warp kernel
            try:
                got = fn(x, t)
            finally:
                lib.zafb_dct_plan_force_direct(plan, 0)
            for c in (0, 17, 36):
                assert_parity(got[c], ofn(x[c], t))
            assert np.array_equal(got, fn(x, t))  # the default route is the same kernel
        assert np.max(np.abs(fn(fn(x, 2), 3) - x)) <= 1e-5
        assert np.max(np.abs(fn(fn(x, 4), 4) - x)) <= 1e-5

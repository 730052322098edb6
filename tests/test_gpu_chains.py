"""Device-resident chains: the reference's own demos are stft -> mask -> istft (zaf.py:162-198)
and mdct -> ... -> imdct (zaf.py:1098-1105).  Here the whole chain runs on DeviceArrays -- nothing crosses PCIe between
the transforms -- and is compared with the oracle doing the same chain in float64 NumPy."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def assert_parity(got, ref, tol=TOL):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    mx, l2 = oracle.parity_metrics(got, ref)
    assert mx <= tol and l2 <= tol, (mx, l2)


def _stereo(seed, ns):
    rng = np.random.default_rng(seed)
    centre = rng.uniform(-0.5, 0.5, ns)
    left = centre + 0.3 * rng.uniform(-1, 1, ns)
    right = centre + 0.3 * rng.uniform(-1, 1, ns)
    return left.astype(np.float32), right.astype(np.float32)


@pytest.mark.parametrize("layout", ["frame_major", "bin_major"])
@pytest.mark.parametrize("n", [2048, 1024])
def test_centre_extraction_chain_on_device(zaf_gpu, layout, n):
    """zaf.py:166-191 with both channels as a batch of two clips: STFT, magnitudes of rows 0..N/2, the two centre masks
    min(|X1|, |X2|) / |Xc|, mask mirrored onto the upper rows, ISTFT."""
    zaf = zaf_gpu
    hop = n // 2
    w = oracle.hamming_periodic(n)
    left, right = _stereo(5, 30000)
    x = np.stack([left, right])
    launches = zaf.launch_count()
    h2d0 = zaf.host_copy_bytes()
    xd = zaf.to_device(x)
    spec = zaf.stft(xd, w, hop, layout=layout)
    mag = zaf.spec_abs(spec, n // 2 + 1)
    assert mag.shape == (2, n // 2 + 1, spec.shape[-1]) and mag.transposed == spec.transposed
    swapped = zaf.DeviceArray(mag.mem_shape, np.float32, transposed=mag.transposed)  # (|X2|, |X1|)
    row = mag.nbytes // 2
    lib, C = zaf._lib.lib(), zaf._lib.C
    zaf._lib.check(lib.zafb_memcpy_d2d(C.c_void_p(swapped.ptr), C.c_void_p(mag.ptr + row), row, None))
    zaf._lib.check(lib.zafb_memcpy_d2d(C.c_void_p(swapped.ptr + row), C.c_void_p(mag.ptr), row, None))
    mask = zaf.ratio_min(mag, swapped)  # clip 0: min(|X1|,|X2|)/|X1|, clip 1: min(|X2|,|X1|)/|X2|
    centre = zaf.spec_mask(spec, mask)
    y = zaf.istft(centre, w, hop).to_host()
    assert zaf.launch_count() - launches >= 5
    assert zaf.host_copy_bytes() == h2d0  # no host-pipeline traffic: the chain stayed on the device

    s1 = oracle.stft(left, w, hop)
    s2 = oracle.stft(right, w, hop)
    a1, a2 = np.abs(s1[: n // 2 + 1]), np.abs(s2[: n // 2 + 1])
    m1, m2 = np.minimum(a1, a2) / a1, np.minimum(a1, a2) / a2
    c1 = np.concatenate((m1, m1[-2:0:-1])) * s1
    c2 = np.concatenate((m2, m2[-2:0:-1])) * s2
    assert_parity(y[0], oracle.istft(c1, w, hop))
    assert_parity(y[1], oracle.istft(c2, w, hop))
    # intermediate stages too
    assert_parity(mag.to_host()[0], a1)
    assert_parity(centre.to_host()[1], c2)


def test_full_mask_and_in_place(zaf_gpu):
    zaf = zaf_gpu
    n, hop = 512, 128
    w = oracle.hamming_periodic(n)
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, 9000).astype(np.float32)
    spec = zaf.stft(zaf.to_device(x), w, hop)
    nt = spec.shape[-1]
    full = rng.uniform(0, 1, (n, nt)).astype(np.float32)
    md = zaf.to_device(np.ascontiguousarray(full.T))
    md = zaf.DeviceArray((nt, n), np.float32, ptr=md.ptr, owner=md, transposed=True)  # frame-major view
    before = spec.to_host().copy()
    out = zaf.spec_mask(spec, md, out=spec)
    assert out is spec
    got = spec.to_host()
    assert np.array_equal(got, before * full)  # one fp32 multiply per component: bit-exact
    with pytest.raises(ValueError):
        zaf.spec_mask(spec, zaf.DeviceArray((nt, 7), np.float32, transposed=True))


def test_mdct_quantise_imdct_chain_on_device(zaf_gpu):
    """mdct -> uniform quantiser -> imdct on the device; the quantiser is bit-exact against NumPy float32 on the same
    coefficients, the synthesis matches the oracle's imdct of those quantised coefficients."""
    zaf = zaf_gpu
    n = 2048
    w = oracle.kbd_window(n)
    rng = np.random.default_rng(23)
    x = rng.uniform(-1, 1, (3, 50000)).astype(np.float32)
    step = np.float32(0.37)
    coef = zaf.mdct(zaf.to_device(x), w)
    q = zaf.quantize(coef, step)
    y = zaf.imdct(q, w).to_host()
    c_host = coef.to_host()
    q_host = q.to_host()
    assert np.array_equal(q_host, step * np.round(c_host / step))
    assert len(np.unique(q_host)) > 50
    for c in range(3):
        assert_parity(c_host[c], oracle.mdct(x[c], w))
        assert_parity(y[c], oracle.imdct(q_host[c].astype(np.float64), w))
    # a real mask on the coefficients (e.g. band-limiting), in place
    keep = np.zeros(c_host.shape, np.float32)
    keep[:, : n // 8, :] = 1.0
    kd = zaf.to_device(np.ascontiguousarray(np.swapaxes(keep, 1, 2)))
    kd = zaf.DeviceArray(kd.mem_shape, np.float32, ptr=kd.ptr, owner=kd, transposed=True)
    zaf.multiply(coef, kd, out=coef)
    assert np.array_equal(coef.to_host(), c_host * keep)


def test_count_mismatch(zaf_gpu):
    zaf = zaf_gpu
    rng = np.random.default_rng(2)
    a = rng.standard_normal(100003).astype(np.float32)
    b = a.copy()
    b[[5, 77, 100002]] += 1.0
    ad, bd = zaf.to_device(a), zaf.to_device(b)
    assert zaf.count_mismatch(ad, ad) == 0
    assert zaf.count_mismatch(ad, bd) == 3


def test_to_host_out_respects_row_padding(zaf_gpu):
    """DeviceArray.to_host(out=...) with padded rows (every batched imdct result has an odd length and an even pitch):
    `out` of the LOGICAL shape receives exactly the logical columns; wrong sizes and dtypes are refused."""
    zaf = zaf_gpu
    w = oracle.kbd_window(512)
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (4, 6000)).astype(np.float32)
    y = zaf.imdct(zaf.mdct(zaf.to_device(x), w), w)
    assert y.cols is not None and y.pitch == y.cols + 1
    want = y.to_host()
    guard = np.full(4 * y.cols + 8, 7.0, np.float32)
    out = guard[: 4 * y.cols].reshape(4, y.cols)
    got = y.to_host(out=out)
    assert np.array_equal(got, want) and np.all(guard[4 * y.cols:] == 7.0)
    with pytest.raises(ValueError):
        y.to_host(out=np.empty((4, y.cols), np.float64))
    with pytest.raises(ValueError):
        y.to_host(out=np.empty((3, y.cols), np.float32))
    with pytest.raises(ValueError):
        zaf.mdct(zaf.to_device(x.astype(np.float64)), w)  # float64 device input is refused, not reinterpreted


@pytest.mark.parametrize("n,hop", [(2048, 512), (1024, 512), (4096, 1024), (512, 64), (256, 64), (2048, 300), (100, 25), (63, 10)])
def test_onesided_stft_istft(zaf_gpu, n, hop):
    """onesided=True (explicit non-reference mode): rows 0 .. N/2 of the reference's spectrum,
    bit-identical to the two-sided result's lower rows on the warp kernels; spec_mirror rebuilds the two-sided spectrum
    bit for bit; the one-sided inverse equals the reference's inverse of the Hermitian-completed spectrum."""
    zaf = zaf_gpu
    rng = np.random.default_rng(n + hop)
    x = rng.uniform(-1, 1, (3, 20000)).astype(np.float32)
    w = oracle.hamming_periodic(n) if n % 2 == 0 else np.hanning(n + 2)[1:-1]
    bins = n // 2 + 1
    full = zaf.stft(x, w, hop)
    half = zaf.stft(x, w, hop, onesided=True)
    assert half.shape == (3, bins, full.shape[-1]) and half.dtype == np.complex64
    assert np.array_equal(half, full[:, :bins])
    for c in range(3):
        assert_parity(half[c], oracle.stft(x[c], w, hop)[:bins])
    xd = zaf.to_device(x)
    hd = zaf.stft(xd, w, hop, onesided=True)
    assert np.array_equal(hd.to_host(), half)
    mirrored = zaf.spec_mirror(hd, n).to_host()
    if n in (256, 512, 1024, 2048, 4096):  # the warp kernels store the upper rows as exact conjugates of the lower ones
        assert np.array_equal(mirrored, full)
    else:                                  # the generic kernels compute every bin on its own
        assert np.array_equal(mirrored[:, :bins], full[:, :bins])
        for c in range(3):
            assert_parity(mirrored[c], oracle.stft(x[c], w, hop))
    # inverse: device and host inputs
    y_full = zaf.istft(full, w, hop)
    y_dev = zaf.istft(hd, w, hop, onesided=True).to_host()
    y_host = zaf.istft(half, w, hop, onesided=True)
    assert y_dev.shape == y_full.shape and np.array_equal(y_dev, y_host)
    for c in range(3):
        ref = oracle.istft(oracle.stft(x[c], w, hop), w, hop)
        assert_parity(y_dev[c], ref)
    # a one-sided spectrum with non-zero imaginary parts in the DC and Nyquist rows: Re(ifft) drops them (zaf.py:223)
    spec = (rng.standard_normal((bins, 40)) + 1j * rng.standard_normal((bins, 40))).astype(np.complex64)
    two = np.concatenate((spec, np.conj(spec[(n - 1) // 2:0:-1])))
    assert two.shape[0] == n
    assert_parity(zaf.istft(spec, w, hop, onesided=True), oracle.istft(two.astype(np.complex128), w, hop))
    with pytest.raises(ValueError):
        zaf.istft(spec[:-1], w, hop, onesided=True)


@pytest.mark.parametrize("n,hop", [(2048, 512), (1024, 256), (2048, 1024)])
@pytest.mark.parametrize("onesided", [False, True])
def test_istft_with_mask_fused_into_the_loads(zaf_gpu, n, hop, onesided):
    """zaf.istft(X, w, hop, mask=M) == istft(np.concatenate((M, M[-2:0:-1])) * X) (zaf.py:185-190): the fused kernel
    (N = 2048, hop = 512) and the multiply-then-transform fallback of every other geometry, two-sided and one-sided
    spectra, a batch and a single clip; also against the unfused device chain."""
    zaf = zaf_gpu
    w = oracle.hamming_periodic(n)
    rng = np.random.default_rng(n + hop)
    x = rng.uniform(-1, 1, (3, 20000)).astype(np.float32)
    xd = zaf.to_device(x)
    spec = zaf.stft(xd, w, hop, onesided=onesided)
    nt = spec.shape[-1]
    m_host = rng.uniform(0, 1, (3, n // 2 + 1, nt)).astype(np.float32)
    mask = zaf.to_device(np.ascontiguousarray(np.swapaxes(m_host, 1, 2)))       # frame-major memory (3, nt, N/2+1)
    mask = zaf.DeviceArray(mask.mem_shape, np.float32, ptr=mask.ptr, owner=mask, transposed=True)
    got = zaf.istft(spec, w, hop, onesided=onesided, mask=mask).to_host()
    unfused = zaf.istft(zaf.spec_mask(spec, mask), w, hop, onesided=onesided).to_host()
    for c in range(3):
        m = m_host[c].astype(np.float64)
        ref = oracle.istft(np.concatenate((m, m[-2:0:-1])) * oracle.stft(x[c], w, hop), w, hop)
        assert_parity(got[c], ref)
        assert_parity(unfused[c], ref)
    one = zaf.stft(zaf.to_device(x[1]), w, hop, onesided=onesided)
    m1_host = np.ascontiguousarray(m_host[1].T)  # (nt, N/2+1): to_device would pad the odd row length like a signal batch
    m1 = zaf.DeviceArray(m1_host.shape, np.float32, transposed=True)
    zaf._lib.check(zaf._lib.lib().zafb_memcpy_h2d(zaf._lib.C.c_void_p(m1.ptr), m1_host.ctypes.data, m1_host.nbytes, None))
    zaf.synchronize()
    y1 = zaf.istft(one, w, hop, onesided=onesided, mask=m1).to_host()
    assert np.max(np.abs(y1 - got[1])) <= 1e-6 * np.max(np.abs(got[1]))
    with pytest.raises(ValueError):
        zaf.istft(spec.to_host(), w, hop, onesided=onesided, mask=mask)
    if not onesided:  # the reference's C-order memory: mask in the same layout, multiply-then-transform
        spec_c = zaf.stft(xd, w, hop, layout="bin_major")
        mask_c = zaf.to_device(m_host.reshape(-1)).to_host().reshape(m_host.shape)  # round trip: plain (3, N/2+1, nt) memory
        md = zaf.DeviceArray(m_host.shape, np.float32)
        zaf._lib.check(zaf._lib.lib().zafb_memcpy_h2d(zaf._lib.C.c_void_p(md.ptr), mask_c.ctypes.data, mask_c.nbytes, None))
        zaf.synchronize()
        yc = zaf.istft(spec_c, w, hop, mask=md).to_host()
        assert np.max(np.abs(yc - got)) <= 1e-5 * np.max(np.abs(got))

"""Seeded random sweeps over the shapes served by the warp-level kernels (window lengths 512 / 1024 / 2048): hop
lengths that do not divide the window, odd hops (generic-kernel fallback), signals shorter than one window, empty
signals, ragged batches -- every result against the float64 oracle, bookkeeping exactly equal."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _check(got, ref, what, scale=1.0):
    """Normalised parity; a reference that is numerically zero (e.g. an istft whose only non-zero sample falls in the
    trimmed region) is compared absolutely against the scale of the input instead."""
    assert got.shape == ref.shape, (got.shape, ref.shape, what)
    if ref.size:
        if np.max(np.abs(ref)) < 1e-9 * scale:
            assert np.max(np.abs(got)) <= 1e-5 * scale, (float(np.max(np.abs(got))), what)
            return
        mx, l2 = oracle.parity_metrics(got, ref)
        assert mx <= TOL and l2 <= TOL, (mx, l2, what)


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_stft_istft_random_shapes(zaf_gpu, n):
    rng = np.random.default_rng(9000 + n)
    w = oracle.hamming_periodic(n)
    for case in range(24):
        hop = int(rng.choice([n // 8, n // 4, n // 2, n, int(rng.integers(1, n + 1)), 2 * int(rng.integers(1, n // 2 + 1))]))
        ns = int(rng.choice([0, 1, n - 1, n, n + 1, int(rng.integers(0, 6 * n))]))
        batch = int(rng.integers(1, 5))
        x = rng.uniform(-1, 1, (batch, ns)).astype(np.float32)
        got = zaf_gpu.stft(x, w, hop)
        for c in range(batch):
            ref = oracle.stft(x[c], w, hop)
            _check(got[c], ref, ("stft", n, hop, ns, batch, c))
        if hop <= n:
            y = zaf_gpu.istft(got, w, hop)
            for c in range(batch):
                _check(y[c], oracle.istft(oracle.stft(x[c], w, hop), w, hop), ("istft", n, hop, ns, batch, c))
        # device-resident path, same kernels
        if ns > 0:
            yd = zaf_gpu.stft(zaf_gpu.to_device(x), w, hop).to_host()
            _check(yd[batch - 1], oracle.stft(x[batch - 1], w, hop), ("stft device", n, hop, ns))


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_mdct_imdct_random_shapes(zaf_gpu, n):
    rng = np.random.default_rng(9100 + n)
    w = oracle.kbd_window(n)
    for case in range(16):
        ns = int(rng.choice([0, 1, n // 2 - 1, n // 2, n, int(rng.integers(0, 8 * n))]))
        batch = int(rng.integers(1, 5))
        x = rng.uniform(-1, 1, (batch, ns)).astype(np.float32)
        got = zaf_gpu.mdct(x, w)
        back = zaf_gpu.imdct(got, w)
        for c in range(batch):
            ref = oracle.mdct(x[c], w)
            _check(got[c], ref, ("mdct", n, ns, batch, c))
            _check(back[c], oracle.imdct(ref, w), ("imdct", n, ns, batch, c))


@pytest.mark.parametrize("n", [512, 1024, 2048])
def test_mel_mfcc_random_shapes(zaf_gpu, n):
    rng = np.random.default_rng(9200 + n)
    w = oracle.hamming_periodic(n)
    for case in range(10):
        fs = int(rng.choice([8000, 16000, 22050, 44100]))
        n_mels = int(rng.integers(2, 129))
        ncoef = int(rng.integers(1, min(n_mels, 64) + 1))
        hop = int(rng.choice([n // 4, n // 2, 2 * int(rng.integers(1, n // 2 + 1))]))
        ns = int(rng.integers(0, 5 * n))
        x = rng.uniform(-1, 1, (2, ns)).astype(np.float32)
        fb = zaf_gpu.melfilterbank(fs, n, n_mels)
        dense = fb.toarray()
        mel = zaf_gpu.melspectrogram(x, w, hop, fb)
        cep = zaf_gpu.mfcc(x, w, hop, fb, ncoef)
        for c in range(2):
            _check(mel[c], oracle.melspectrogram(x[c], w, hop, dense), ("mel", n, fs, n_mels, hop, ns, c))
            ref = oracle.mfcc(x[c], w, hop, dense, ncoef)
            assert cep[c].shape == ref.shape
            if ref.size and np.max(np.abs(ref)) > 0:
                _check(cep[c], ref, ("mfcc", n, fs, n_mels, ncoef, hop, ns, c))


def test_dct_dst_random_batches(zaf_gpu):
    rng = np.random.default_rng(9300)
    for case in range(12):
        n = int(rng.choice([1024, 1024, 512, 256, 1000, 37]))
        batch = int(rng.integers(1, 40))
        x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
        for fn, ofn in ((zaf_gpu.dct, oracle.dct), (zaf_gpu.dst, oracle.dst)):
            for t in (1, 2, 3, 4):
                got = fn(x, t)
                for c in (0, batch - 1):
                    _check(got[c], ofn(x[c], t), (fn.__name__, t, n, batch, c))

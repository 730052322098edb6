#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/zaf.py (with an empty ``matplotlib`` stub, the
reference's only missing import -- the hot path never touches it), feeds it
small float64 inputs and stores inputs + reference outputs as compressed
``.npz`` files.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4); these files are the pin for ``oracle/`` and for the CUDA
kernels.  Inputs are stored as float32-representable float64 values so the fp32
GPU path sees exactly the same numbers.
"""
import os
import sys
import types
import wave

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import zaf  # noqa: E402

    return zaf


def wav_mono():
    with wave.open(os.path.join(REF, "audio_file.wav")) as f:
        raw = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16).reshape(-1, f.getnchannels())
    return (raw / 2.0 ** 15).mean(axis=1), 44100  # same normalisation as zaf.wavread (zaf.py:1202)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def hamming_p(n):
    import scipy.signal

    return f32(scipy.signal.windows.hamming(n, sym=False))


def main():
    zaf = load_reference()
    rng = np.random.default_rng(20261017)
    wav, fs = wav_mono()
    out = {}

    # ---- stft / istft ---------------------------------------------------------------
    stft_cases = {
        "wav_n2048_h512": (f32(wav[100000:108000]), hamming_p(2048), 512),
        "wav_n2048_h1024": (f32(wav[200000:206000]), hamming_p(2048), 1024),
        "rand_n256_h64": (f32(rng.uniform(-1, 1, 1000)), hamming_p(256), 64),
        "rand_n256_h100": (f32(rng.uniform(-1, 1, 1000)), hamming_p(256), 100),
        "rand_n1024_h256": (f32(rng.uniform(-1, 1, 5000)), hamming_p(1024), 256),
        "empty_n64_h16": (np.zeros(0), hamming_p(64), 16),
        "short_n128_h32": (f32(rng.uniform(-1, 1, 37)), hamming_p(128), 32),
        "odd_n63_h10": (f32(rng.uniform(-1, 1, 200)), f32(np.hanning(63)), 10),
        "hop_gt_half_n64_h48": (f32(rng.uniform(-1, 1, 500)), hamming_p(64), 48),
    }
    g = {}
    for name, (x, w, hop) in stft_cases.items():
        spec = zaf.stft(x, w, hop)
        g[f"{name}/x"] = x
        g[f"{name}/w"] = w
        g[f"{name}/hop"] = np.int64(hop)
        g[f"{name}/stft"] = spec
        g[f"{name}/istft"] = zaf.istft(spec, w, hop)
    # istft of a non-Hermitian spectrum (Re(ifft) discards the rest, zaf.py:223)
    spec = rng.standard_normal((128, 9)) + 1j * rng.standard_normal((128, 9))
    spec = spec.astype(np.complex64).astype(np.complex128)
    g["nonherm_n128_h32/spec"] = spec
    g["nonherm_n128_h32/w"] = hamming_p(128)
    g["nonherm_n128_h32/hop"] = np.int64(32)
    g["nonherm_n128_h32/istft"] = zaf.istft(spec, hamming_p(128), 32)
    out["stft"] = g

    # ---- integer bookkeeping sweep -------------------------------------------------
    rows = []
    for n in (8, 63, 64, 256, 2048):
        for hop in (1, 3, n // 4, n // 2, n // 2 + 1, n):
            hop = max(hop, 1)
            for ns in (0, 1, n - 1, n, n + 1, 1000, 4801, 48000):
                spec = zaf.stft(np.zeros(ns), np.ones(n), hop)
                rows.append((ns, n, hop, spec.shape[1], len(zaf.istft(spec, np.ones(n), hop))))
    mrows = []
    for n in (4, 64, 256, 2048):
        for ns in (0, 1, n // 2 - 1, n // 2, n // 2 + 1, n, 1000, 44100):
            m = zaf.mdct(np.zeros(ns), np.ones(n))
            mrows.append((ns, n, m.shape[1], len(zaf.imdct(m, np.ones(n)))))
    out["geometry"] = {"stft_rows": np.array(rows, dtype=np.int64), "mdct_rows": np.array(mrows, dtype=np.int64)}

    # ---- mel filterbank / melspectrogram / mfcc ------------------------------------
    g = {}
    g["fb_16k_1024_128"] = zaf.melfilterbank(16000, 1024, 128).toarray()
    g["fb_44k_2048_128"] = zaf.melfilterbank(44100, 2048, 128).toarray()
    g["fb_8k_256_20"] = zaf.melfilterbank(8000, 256, 20).toarray()
    x = f32(wav[300000:312000])
    w = hamming_p(2048)
    fb = zaf.melfilterbank(fs, 2048, 128)
    g["wav/x"], g["wav/w"], g["wav/hop"] = x, w, np.int64(1024)
    g["wav/fs"], g["wav/nmel"], g["wav/ncoef"] = np.int64(fs), np.int64(128), np.int64(20)
    g["wav/mel"] = zaf.melspectrogram(x, w, 1024, fb)
    g["wav/mfcc"] = zaf.mfcc(x, w, 1024, fb, 20)
    x = f32(rng.uniform(-1, 1, 6000))
    w = hamming_p(1024)
    fb = zaf.melfilterbank(16000, 1024, 128)
    g["cfg3/x"], g["cfg3/w"], g["cfg3/hop"] = x, w, np.int64(256)
    g["cfg3/fs"], g["cfg3/nmel"], g["cfg3/ncoef"] = np.int64(16000), np.int64(128), np.int64(40)
    g["cfg3/mel"] = zaf.melspectrogram(x, w, 256, fb)
    g["cfg3/mfcc"] = zaf.mfcc(x, w, 256, fb, 40)
    # a silent clip: every kept cepstral coefficient must be 0 (log(eps) is constant)
    x = np.zeros(3000)
    g["silent/x"], g["silent/w"], g["silent/hop"] = x, w, np.int64(256)
    g["silent/fs"], g["silent/nmel"], g["silent/ncoef"] = np.int64(16000), np.int64(128), np.int64(40)
    g["silent/mel"] = zaf.melspectrogram(x, w, 256, fb)
    g["silent/mfcc"] = zaf.mfcc(x, w, 256, fb, 40)
    out["mel"] = g

    # ---- CQT ------------------------------------------------------------------------
    g = {}
    for tag, (res, fmin, fmax) in {
        "c1c8_12": (12, 32.70319566257483, 4186.009044809578),
        "a1a7_24": (24, 55.0, 3520.0),
    }.items():
        k = scipy.sparse.csr_matrix(zaf.cqtkernel(fs, res, fmin, fmax))
        g[f"{tag}/kernel_data"] = k.data
        g[f"{tag}/kernel_indices"] = k.indices.astype(np.int32)
        g[f"{tag}/kernel_indptr"] = k.indptr.astype(np.int32)
        g[f"{tag}/kernel_shape"] = np.array(k.shape, dtype=np.int64)
        g[f"{tag}/params"] = np.array([fs, res, fmin, fmax], dtype=np.float64)
        x = f32(wav[400000:400000 + 17640])  # 0.4 s -> 10 frames at 25 fps
        g[f"{tag}/x"] = x
        g[f"{tag}/time_resolution"] = np.int64(25)
        g[f"{tag}/spec"] = zaf.cqtspectrogram(x, fs, 25, k)
        g[f"{tag}/chroma"] = zaf.cqtchromagram(x, fs, 25, res, k)
    out["cqt"] = g

    # ---- DCT / DST ------------------------------------------------------------------
    g = {}
    vecs = {"wav1024": f32(wav[500000:501024]), "rand64": f32(rng.uniform(-1, 1, 64)),
            "rand8": f32(rng.uniform(-1, 1, 8)), "rand2048": f32(rng.uniform(-1, 1, 2048)),
            "rand100": f32(rng.uniform(-1, 1, 100))}
    for name, v in vecs.items():
        g[f"{name}/x"] = v
        for t in (1, 2, 3, 4):
            g[f"{name}/dct{t}"] = zaf.dct(v, t)
            g[f"{name}/dst{t}"] = zaf.dst(v, t)
    out["dctdst"] = g

    # ---- MDCT / IMDCT ---------------------------------------------------------------
    g = {}
    kbd = np.kaiser(1025, 5 * np.pi)
    kbd_half = np.sqrt(np.cumsum(kbd[:1024]) / np.sum(kbd))
    cases = {
        "wav_kbd2048": (f32(wav[600000:608192]), f32(np.concatenate([kbd_half, kbd_half[::-1]]))),
        "rand_sine256": (f32(rng.uniform(-1, 1, 1000)), f32(np.sin(np.pi / 256 * (np.arange(256) + 0.5)))),
        "rand_sine64_exact": (f32(rng.uniform(-1, 1, 320)), f32(np.sin(np.pi / 64 * (np.arange(64) + 0.5)))),
        "empty_sine64": (np.zeros(0), f32(np.sin(np.pi / 64 * (np.arange(64) + 0.5)))),
        "rand_sine1024": (f32(rng.uniform(-1, 1, 3000)), f32(np.sin(np.pi / 1024 * (np.arange(1024) + 0.5)))),
    }
    for name, (x, w) in cases.items():
        m = zaf.mdct(x, w)
        g[f"{name}/x"], g[f"{name}/w"] = x, w
        g[f"{name}/mdct"] = m
        g[f"{name}/imdct"] = zaf.imdct(m, w)
    out["mdct"] = g

    for fname, arrays in out.items():
        path = os.path.join(HERE, f"golden_{fname}.npz")
        np.savez_compressed(path, **arrays)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays")


if __name__ == "__main__":
    main()

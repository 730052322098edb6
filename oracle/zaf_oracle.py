"""NumPy float64 restatement of the transform hot path of the reference ``zaf.py``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Never imported by the
product.

Two families of functions live here:

* ``stft, istft, melspectrogram, mfcc, cqtspectrogram, cqtchromagram, dct, dst,
  mdct, imdct, melfilterbank, cqtkernel`` -- the *port*: the same sequence of
  operations the reference performs (same framing loop, same FFT sizes, same
  library calls), restated in this repository's own words.  Because the work
  per call is the same as the reference's, this is also what ``bench.py``
  times as the CPU baseline (``cpu_baseline.kind == "port"``).
* ``*_direct`` -- closed-form O(N^2) definitions (explicit DFT / cosine / sine
  matrices) that do not go through an FFT library at all.  They pin the port
  (and therefore pocketfft) independently for small sizes.

The arithmetic proper lives in third-party code that is *not* vendored in the
reference and that the reference does not pin (README.md:18): NumPy
``np.fft.fft/ifft`` (pocketfft), SciPy ``scipy.fftpack.dct``, ``np.matmul``
(OpenBLAS) and SciPy CSR mat-vec.  De-facto pins are the versions in this
image: NumPy 2.3.5, SciPy 1.18.1.

All inputs are up-cast to float64 first: the reference computes in float64 for
every path except a few NumPy>=2 dtype leaks on float32 input (SURVEY.md
section 8c), and float64 is the contract the fp32 GPU kernels are compared with.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "stft_geometry", "istft_length", "mdct_geometry", "imdct_length", "cqt_geometry",
    "stft", "istft", "melfilterbank", "melspectrogram", "mfcc", "cqtkernel",
    "cqtspectrogram", "cqtchromagram", "dct", "dst", "mdct", "imdct",
    "stft_direct", "istft_direct", "dct_direct", "dst_direct", "mdct_direct",
    "imdct_direct", "dct2_ortho_matrix", "hamming_periodic", "kbd_window",
    "sine_window", "parity_metrics",
]


# ----------------------------------------------------------------------------
# Integer bookkeeping (must be bit-exact).  Plain Python ints.
# ----------------------------------------------------------------------------
def stft_geometry(number_samples: int, window_length: int, step_length: int):
    """(front pad, number of frames, tail pad) of ``zaf.stft`` -- zaf.py:95-125.

    The reference uses a *float* division followed by ``np.ceil`` (zaf.py:102-109);
    that is what is reproduced here.
    """
    pad = int(math.floor(window_length / 2))                                   # :99
    nt = int(math.ceil(((number_samples + 2 * pad) - window_length) / step_length)) + 1  # :102-109
    tail = (nt * step_length + (window_length - step_length) - pad) - number_samples     # :116-121
    return pad, nt, tail


def istft_length(window_length: int, number_times: int, step_length: int):
    """(overlap-add length, trim at each end, output length) -- zaf.py:217, 236-238."""
    total = number_times * step_length + (window_length - step_length)         # :217
    trim = window_length - step_length                                         # :236-238
    return total, trim, len(range(total)[trim:total - trim])                   # Python slice semantics


def mdct_geometry(number_samples: int, window_length: int):
    """(M, number of frames, front pad, tail pad) of ``zaf.mdct`` -- zaf.py:1029-1041."""
    half = int(window_length / 2)                                              # :1029-1030
    nt = int(math.ceil(number_samples / half)) + 1                             # :1033
    return half, nt, half, (nt + 1) * half - number_samples                    # :1036-1041


def imdct_length(number_frequencies: int, number_times: int):
    """(overlap-add length, output length) of ``zaf.imdct`` -- zaf.py:1132, 1182."""
    total = number_frequencies * (number_times + 1)                            # :1132
    # slice [M : -M-1] of an array of length M*(nt+1)                            :1182
    return total, max(total - 2 * number_frequencies - 1, 0)


def cqt_geometry(number_samples: int, sampling_frequency, time_resolution, fft_length: int):
    """(step, number of frames, front pad, tail pad) -- zaf.py:603-620.

    ``round`` is Python's round-half-to-even on a float, exactly as in the reference.
    """
    step = round(sampling_frequency / time_resolution)                         # :603
    nt = int(math.floor(number_samples / step))                                # :606
    front = int(math.ceil((fft_length - step) / 2))                            # :615
    back = int(math.floor((fft_length - step) / 2))                            # :616
    return step, nt, front, back


# ----------------------------------------------------------------------------
# Windows used by the tests / bench (not part of zaf.py itself)
# ----------------------------------------------------------------------------
def hamming_periodic(n: int) -> np.ndarray:
    """Periodic Hamming window == scipy.signal.windows.hamming(n, sym=False)."""
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)


def sine_window(n: int) -> np.ndarray:
    """Sine window; satisfies Princen-Bradley (used with mdct/imdct)."""
    return np.sin(np.pi / n * (np.arange(n) + 0.5))


def kbd_window(n: int, alpha: float = 5.0) -> np.ndarray:
    """Proper Kaiser-Bessel-derived window of length n."""
    k = np.kaiser(n // 2 + 1, alpha * np.pi)
    half = np.sqrt(np.cumsum(k[: n // 2]) / np.sum(k))
    return np.concatenate([half, half[::-1]])


# ----------------------------------------------------------------------------
# The port
# ----------------------------------------------------------------------------
def _f64(a):
    return np.asarray(a, dtype=np.float64)


def stft(audio_signal, window_function, step_length):
    """Two-sided STFT, shape (window_length, number_times), complex128 -- zaf.py:95-141."""
    x = _f64(audio_signal)
    w = _f64(window_function)
    if x.ndim != 1:
        raise ValueError("audio_signal must be 1-D (zaf.py:132-136 broadcasts otherwise)")
    n = len(w)
    pad, nt, tail = stft_geometry(len(x), n, step_length)
    padded = np.concatenate((np.zeros(pad), x, np.zeros(tail)))                # :112-125
    frames = np.zeros((n, nt))                                                 # :128
    start = 0
    for j in range(nt):                                                        # :131-136
        frames[:, j] = padded[start:start + n] * w
        start += step_length
    return np.fft.fft(frames, axis=0)                                          # :139


def istft(audio_stft, window_function, step_length):
    """Inverse STFT by overlap-add -- zaf.py:214-243 (no synthesis window, COLA gain only)."""
    spec = np.asarray(audio_stft, dtype=np.complex128)
    w = _f64(window_function)
    n, nt = spec.shape                                                         # :214
    total, trim, _ = istft_length(n, nt, step_length)
    frames = np.fft.ifft(spec, axis=0).real                                    # :223
    y = np.zeros(total)                                                        # :220
    start = 0
    for j in range(nt):                                                        # :227-233
        y[start:start + n] += frames[:, j]
        start += step_length
    y = y[trim:total - trim]                                                   # :236-238
    gain = sum(w[0:n:step_length])                                             # :241 builtin sum
    return y / gain


def melfilterbank(sampling_frequency, window_length, number_filters):
    """Dense (number_filters, window_length/2) mel filterbank -- zaf.py:280-316.

    The reference returns ``scipy.sparse.csr_matrix`` of the same values (:319);
    the oracle keeps it dense (``melspectrogram`` densifies it anyway, :373).
    """
    mel_lo = 2595 * np.log10(1 + (sampling_frequency / window_length) / 700)   # :280
    mel_hi = 2595 * np.log10(1 + (sampling_frequency / 2) / 700)               # :281
    width = 2 * (mel_hi - mel_lo) / (number_filters + 1)                       # :284
    mel_points = np.arange(mel_lo, mel_hi + 1, width / 2)                      # :287
    idx = np.round(700 * (np.power(10, mel_points / 2595) - 1)
                   * window_length / sampling_frequency).astype(int)           # :290-295
    bank = np.zeros((number_filters, int(window_length / 2)))                  # :298
    for i in range(number_filters):                                            # :301-316
        lo, mid, hi = idx[i], idx[i + 1], idx[i + 2]
        bank[i, lo - 1:mid] = np.linspace(0, 1, num=mid - lo + 1)
        bank[i, mid - 1:hi] = np.linspace(1, 0, num=hi - mid + 1)
    return bank


def _dense(op):
    return op.toarray() if hasattr(op, "toarray") else np.asarray(op)


def melspectrogram(audio_signal, window_function, step_length, mel_filterbank):
    """(number_mels, number_times) -- zaf.py:369-375 (rows 1..N/2 of the STFT: no DC, with Nyquist)."""
    spec = stft(audio_signal, window_function, step_length)                    # :369
    mag = np.abs(spec[1:int(len(window_function) / 2) + 1, :])                 # :370
    return np.matmul(_dense(mel_filterbank), mag)                              # :373


def dct2_ortho_matrix(n: int) -> np.ndarray:
    """Orthonormal DCT-II matrix D (D @ v == scipy.fftpack.dct(v, norm='ortho'))."""
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    d = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * m + 1) * k / (2 * n))
    d[0, :] /= np.sqrt(2.0)
    return d


def mfcc(audio_signal, window_function, step_length, mel_filterbank, number_coefficients):
    """(number_coefficients, number_times) -- zaf.py:436-454."""
    import scipy.fftpack

    spec = stft(audio_signal, window_function, step_length)                    # :436
    power = np.abs(spec[1:int(len(window_function) / 2) + 1, :]) ** 2          # :437-439
    logmel = np.log(np.matmul(_dense(mel_filterbank), power) + np.finfo(float).eps)  # :444-446
    cep = scipy.fftpack.dct(logmel, axis=0, norm="ortho")                      # :443-449
    return cep[1:number_coefficients + 1, :]                                   # :452


def cqtkernel(sampling_frequency, octave_resolution, minimum_frequency, maximum_frequency):
    """Dense complex (number_frequencies, fft_length) CQT kernel -- zaf.py:497-559.

    The reference returns the CSR form of the same values (:554-557).
    """
    q = 1 / (pow(2, 1 / octave_resolution) - 1)                                # :497
    nf = round(octave_resolution * np.log2(maximum_frequency / minimum_frequency))   # :500-502
    fft_length = int(pow(2, np.ceil(np.log2(q * sampling_frequency / minimum_frequency))))  # :505-509
    kern = np.zeros((nf, fft_length), dtype=complex)                           # :512
    for i in range(nf):                                                        # :515-544
        f = minimum_frequency * pow(2, i / octave_resolution)
        wl = 2 * round(q * sampling_frequency / f / 2) + 1
        n = np.arange(-(wl - 1) / 2, (wl - 1) / 2 + 1)
        temporal = np.hamming(wl) * np.exp(2 * np.pi * 1j * q * n / wl) / wl
        off = int((fft_length - wl + 1) / 2)
        kern[i, off:off + wl] = temporal
    kern = np.fft.fft(kern, axis=1)                                            # :548
    kern[np.absolute(kern) < 0.01] = 0                                         # :551
    return np.conjugate(kern) / fft_length                                     # :557


def cqtspectrogram(audio_signal, sampling_frequency, time_resolution, cqt_kernel):
    """(number_frequencies, number_times) magnitude CQT -- zaf.py:603-635."""
    x = _f64(audio_signal)
    nf, fft_length = np.shape(cqt_kernel)                                      # :609
    step, nt, front, back = cqt_geometry(len(x), sampling_frequency, time_resolution, fft_length)
    padded = np.concatenate((np.zeros(front), x, np.zeros(back)))              # :612-620
    out = np.zeros((nf, nt))                                                   # :623
    start = 0
    for j in range(nt):                                                        # :627-633
        spectrum = np.fft.fft(padded[start:start + fft_length])
        out[:, j] = np.absolute(cqt_kernel @ spectrum)                         # CSR `*` == mat-vec
        start += step
    return out


def cqtchromagram(audio_signal, sampling_frequency, time_resolution, octave_resolution, cqt_kernel):
    """(octave_resolution, number_times) -- zaf.py:682-700 (fold rows i::octave_resolution)."""
    spec = cqtspectrogram(audio_signal, sampling_frequency, time_resolution, cqt_kernel)
    nf, nt = spec.shape
    chroma = np.zeros((octave_resolution, nt))
    for i in range(octave_resolution):                                         # :693-698
        chroma[i, :] = spec[i:nf:octave_resolution, :].sum(axis=0)
    return chroma


def dct(audio_signal, dct_type):
    """Orthonormal DCT-I..IV of one vector through a mirrored/zero-stuffed FFT -- zaf.py:759-839.

    Unknown types fall through and return None, like the reference.
    """
    x = _f64(audio_signal).copy()
    n = len(x)
    if dct_type == 1:                                                          # :758-777
        x[[0, -1]] *= np.sqrt(2)
        ext = np.concatenate((x, x[-2:0:-1]))                                  # length 2(n-1)
        out = np.fft.fft(ext).real[:n] / 2
        out[[0, -1]] /= np.sqrt(2)
        return out * np.sqrt(2 / (n - 1))
    if dct_type == 2:                                                          # :779-795
        ext = np.zeros(4 * n)
        ext[1:2 * n:2] = x
        ext[2 * n + 1:4 * n:2] = x[::-1]
        out = np.fft.fft(ext).real[:n] / 2
        out[0] /= np.sqrt(2)
        return out * np.sqrt(2 / n)
    if dct_type == 3:                                                          # :797-819
        x[0] *= np.sqrt(2)
        ext = np.zeros(4 * n)
        ext[0:n] = x
        ext[n + 1:2 * n + 1] = -x[::-1]
        ext[2 * n + 1:3 * n] = -x[1:]
        ext[3 * n + 1:4 * n] = x[:0:-1]
        out = np.fft.fft(ext).real[1:2 * n:2] / 4
        return out * np.sqrt(2 / n)
    if dct_type == 4:                                                          # :821-839
        ext = np.zeros(8 * n)
        ext[1:2 * n:2] = x
        ext[2 * n + 1:4 * n:2] = -x[::-1]
        ext[4 * n + 1:6 * n:2] = -x
        ext[6 * n + 1:8 * n:2] = x[::-1]
        out = np.fft.fft(ext).real[1:2 * n:2] / 4
        return out * np.sqrt(2 / n)
    return None


def dst(audio_signal, dst_type):
    """Orthonormal DST-I..IV of one vector -- zaf.py:901-981."""
    x = _f64(audio_signal).copy()
    n = len(x)
    if dst_type == 1:                                                          # :902-917
        ext = np.zeros(2 * n + 2)
        ext[1:n + 1] = x
        ext[n + 2:] = -x[::-1]
        out = -np.fft.fft(ext).imag[1:n + 1] / 2
        return out * np.sqrt(2 / (n + 1))
    if dst_type == 2:                                                          # :919-935
        ext = np.zeros(4 * n)
        ext[1:2 * n:2] = x
        ext[2 * n + 1:4 * n:2] = -x[::-1]
        out = -np.fft.fft(ext).imag[1:n + 1] / 2
        out[-1] /= np.sqrt(2)
        return out * np.sqrt(2 / n)
    if dst_type == 3:                                                          # :937-959
        x[-1] *= np.sqrt(2)
        ext = np.zeros(4 * n)
        ext[1:n + 1] = x
        ext[n + 1:2 * n] = x[-2::-1]
        ext[2 * n + 1:3 * n + 1] = -x
        ext[3 * n + 1:4 * n] = -x[-2::-1]
        out = -np.fft.fft(ext).imag[1:2 * n:2] / 4
        return out * np.sqrt(2 / n)
    if dst_type == 4:                                                          # :961-981
        ext = np.zeros(8 * n)
        ext[1:2 * n:2] = x
        ext[2 * n + 1:4 * n:2] = x[::-1]
        ext[4 * n + 1:6 * n:2] = -x
        ext[6 * n + 1:8 * n:2] = -x[::-1]
        out = -np.fft.fft(ext).imag[1:2 * n:2] / 4
        return out * np.sqrt(2 / n)
    return None


def mdct(audio_signal, window_function):
    """(window_length/2, number_times) MDCT, unscaled cosine kernel -- zaf.py:1025-1075."""
    x = _f64(audio_signal)
    w = _f64(window_function)
    n = len(w)
    if n % 2:
        raise ValueError("window_function must have an even length (zaf.py:1071 broadcasts otherwise)")
    half, nt, front, tail = mdct_geometry(len(x), n)
    padded = np.concatenate((np.zeros(front), x, np.zeros(tail)))              # :1036-1041
    pre = np.exp(-1j * np.pi / n * np.arange(0, n))                            # :1047-1049
    post = np.exp(-1j * np.pi / n * (n / 2 + 1) * np.arange(0.5, n / 2 + 0.5))  # :1050-1056
    out = np.zeros((half, nt))                                                 # :1044
    start = 0
    for j in range(nt):                                                        # :1061-1073
        seg = np.fft.fft(padded[start:start + n] * w * pre)
        out[:, j] = (seg[:half] * post).real
        start += half
    return out


def imdct(audio_mdct, window_function):
    """Inverse MDCT with TDAC overlap-add -- zaf.py:1125-1184."""
    spec = _f64(audio_mdct)
    w = _f64(window_function)
    half, nt = spec.shape                                                      # :1126
    n = 2 * half
    total, _ = imdct_length(half, nt)
    pre = np.exp(-1j * np.pi / n * (half + 1) * np.arange(0, half))            # :1138-1144
    post = np.exp(-1j * np.pi / n * np.arange(0.5 + half / 2, n + half / 2 + 0.5)) / half  # :1145-1156
    frames = np.fft.fft(spec * pre[:, None], n=n, axis=0)                      # :1159-1163
    frames = 2 * ((frames * post[:, None]).real * w[:, None])                  # :1166-1169
    y = np.zeros(total)                                                        # :1135
    start = 0
    for j in range(nt):                                                        # :1173-1179
        y[start:start + n] += frames[:, j]
        start += half
    return y[half:-half - 1]                                                   # :1182


# ----------------------------------------------------------------------------
# Closed forms (no FFT library): pins for small sizes
# ----------------------------------------------------------------------------
def stft_direct(audio_signal, window_function, step_length):
    """X[k,j] = sum_n w[n] x~[j*hop+n] exp(-2 pi i k n / N) with an explicit DFT matrix."""
    x = _f64(audio_signal)
    w = _f64(window_function)
    n = len(w)
    pad, nt, tail = stft_geometry(len(x), n, step_length)
    padded = np.concatenate((np.zeros(pad), x, np.zeros(tail)))
    idx = np.arange(nt)[None, :] * step_length + np.arange(n)[:, None]
    frames = padded[idx] * w[:, None]
    kn = np.outer(np.arange(n), np.arange(n)) % n
    dft = np.exp(-2j * np.pi * kn / n)
    return dft @ frames


def istft_direct(audio_stft, window_function, step_length):
    """Gather-form overlap-add of Re(IDFT) with an explicit inverse DFT matrix."""
    spec = np.asarray(audio_stft, dtype=np.complex128)
    w = _f64(window_function)
    n, nt = spec.shape
    total, trim, out_len = istft_length(n, nt, step_length)
    kn = np.outer(np.arange(n), np.arange(n)) % n
    frames = ((np.exp(2j * np.pi * kn / n) @ spec) / n).real
    y = np.zeros(out_len)
    for m in range(out_len):
        p = m + trim
        j_lo = max(0, -((n - 1 - p) // step_length))        # ceil((p-n+1)/hop)
        j_hi = min(nt - 1, p // step_length)
        for j in range(j_lo, j_hi + 1):
            y[m] += frames[p - j * step_length, j]
    gain = 0.0
    for v in w[0:n:step_length]:
        gain += v
    return y / gain


def dct_direct(audio_signal, dct_type):
    """Orthonormal DCT matrices, SURVEY.md section 8(a) row a8."""
    x = _f64(audio_signal)
    n = len(x)
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    if dct_type == 1:
        a = np.ones(n)
        a[[0, -1]] = 1 / np.sqrt(2)
        mat = np.sqrt(2 / (n - 1)) * a[:, None] * a[None, :] * np.cos(np.pi * m * k / (n - 1))
    elif dct_type == 2:
        mat = dct2_ortho_matrix(n)
    elif dct_type == 3:
        mat = dct2_ortho_matrix(n).T
    elif dct_type == 4:
        mat = np.sqrt(2 / n) * np.cos(np.pi * (2 * m + 1) * (2 * k + 1) / (4 * n))
    else:
        return None
    return mat @ x


def dst_direct(audio_signal, dst_type):
    """Orthonormal DST matrices, SURVEY.md section 8(a) row a9."""
    x = _f64(audio_signal)
    n = len(x)
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    if dst_type == 1:
        mat = np.sqrt(2 / (n + 1)) * np.sin(np.pi * (m + 1) * (k + 1) / (n + 1))
    elif dst_type in (2, 3):
        c = np.ones(n)
        c[-1] = 1 / np.sqrt(2)
        mat = np.sqrt(2 / n) * c[:, None] * np.sin(np.pi * (2 * m + 1) * (k + 1) / (2 * n))
        if dst_type == 3:
            mat = mat.T
    elif dst_type == 4:
        mat = np.sqrt(2 / n) * np.sin(np.pi * (2 * m + 1) * (2 * k + 1) / (4 * n))
    else:
        return None
    return mat @ x


def _mdct_cos(half: int) -> np.ndarray:
    n = np.arange(2 * half)[None, :]
    k = np.arange(half)[:, None]
    return np.cos(np.pi / half * (n + 0.5 + half / 2) * (k + 0.5))


def mdct_direct(audio_signal, window_function):
    """X[k,j] = sum_n w[n] x~[jM+n] cos(pi/M (n+1/2+M/2)(k+1/2)) -- SURVEY.md 8(a) row a6."""
    x = _f64(audio_signal)
    w = _f64(window_function)
    half, nt, front, tail = mdct_geometry(len(x), len(w))
    padded = np.concatenate((np.zeros(front), x, np.zeros(tail)))
    idx = np.arange(nt)[None, :] * half + np.arange(2 * half)[:, None]
    return _mdct_cos(half) @ (padded[idx] * w[:, None])


def imdct_direct(audio_mdct, window_function):
    """y = trim(OLA_M((2/M) w[n] sum_k X[k,j] cos(...))) -- SURVEY.md 8(a) row a7."""
    spec = _f64(audio_mdct)
    w = _f64(window_function)
    half, nt = spec.shape
    total, _ = imdct_length(half, nt)
    frames = (2.0 / half) * w[:, None] * (_mdct_cos(half).T @ spec)
    y = np.zeros(total)
    for j in range(nt):
        y[j * half:j * half + 2 * half] += frames[:, j]
    return y[half:-half - 1]


# ----------------------------------------------------------------------------
# Parity metric: normalised max-abs and relative L2.
# ----------------------------------------------------------------------------
def parity_metrics(got, ref):
    got = np.asarray(got)
    ref = np.asarray(ref)
    if got.shape != ref.shape:
        raise AssertionError(f"shape mismatch: got {got.shape}, reference {ref.shape}")
    if ref.size == 0:
        return 0.0, 0.0
    diff = np.abs(got.astype(np.complex128) - ref.astype(np.complex128))
    scale = float(np.max(np.abs(ref)))
    norm = float(np.linalg.norm(ref.ravel()))
    max_rel = float(diff.max()) / scale if scale > 0 else float(diff.max())
    l2_rel = float(np.linalg.norm(diff.ravel())) / norm if norm > 0 else float(np.linalg.norm(diff.ravel()))
    return max_rel, l2_rel

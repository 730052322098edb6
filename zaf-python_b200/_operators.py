"""Host-side, one-time operator constructors consumed by the hot path.

They are tiny (milliseconds, once per configuration), so they stay on the host in float64 NumPy
and reproduce the reference's arithmetic step for step so that the operators are bit-identical;
what the GPU consumes is their packed banded form, built at plan creation in csrc/.
"""
from __future__ import annotations

import numpy as np


def melfilterbank(sampling_frequency, window_length, number_filters):
    """Mel filterbank, ``scipy.sparse.csr_matrix`` of shape (number_filters, window_length/2) --
    same construction as ``zaf.melfilterbank`` (zaf.py:246-321): triangular filters between
    rounded mel-spaced bin indices, left slope 0->1 then right slope 1->0 (written second)."""
    import scipy.sparse

    fs, n = sampling_frequency, window_length
    lowest = 2595 * np.log10(1 + (fs / n) / 700)                       # zaf.py:280
    highest = 2595 * np.log10(1 + (fs / 2) / 700)                      # zaf.py:281
    spacing = 2 * (highest - lowest) / (number_filters + 1)            # zaf.py:284
    mels = np.arange(lowest, highest + 1, spacing / 2)                 # zaf.py:287
    edges = np.round(700 * (np.power(10, mels / 2595) - 1) * n / fs).astype(int)   # zaf.py:290-295
    bank = np.zeros((number_filters, int(n / 2)))
    for row in range(number_filters):                                  # zaf.py:301-316
        left, centre, right = (int(e) for e in edges[row:row + 3])
        bank[row, left - 1:centre] = np.linspace(0, 1, num=centre - left + 1)
        bank[row, centre - 1:right] = np.linspace(1, 0, num=right - centre + 1)
    return scipy.sparse.csr_matrix(bank)                               # zaf.py:319


def cqtkernel(sampling_frequency, octave_resolution, minimum_frequency, maximum_frequency):
    """CQT kernel, ``scipy.sparse.csr_matrix`` of shape (number_frequencies, fft_length), complex --
    same construction as ``zaf.cqtkernel`` (zaf.py:457-559): centred Hamming-windowed complex
    exponentials, FFT along rows, magnitudes < 0.01 dropped, conjugated and divided by fft_length."""
    import scipy.sparse

    fs = sampling_frequency
    quality = 1 / (pow(2, 1 / octave_resolution) - 1)                                       # zaf.py:497
    rows = round(octave_resolution * np.log2(maximum_frequency / minimum_frequency))        # zaf.py:500-502
    length = int(pow(2, np.ceil(np.log2(quality * fs / minimum_frequency))))                # zaf.py:505-509
    kernel = np.zeros((rows, length), dtype=complex)
    for row in range(rows):                                                                 # zaf.py:515-544
        frequency = minimum_frequency * pow(2, row / octave_resolution)
        width = 2 * round(quality * fs / frequency / 2) + 1
        positions = np.arange(-(width - 1) / 2, (width - 1) / 2 + 1)
        atom = np.hamming(width) * np.exp(2 * np.pi * 1j * quality * positions / width) / width
        first = int((length - width + 1) / 2)
        kernel[row, first:first + width] = atom
    kernel = np.fft.fft(kernel, axis=1)                                                     # zaf.py:548
    kernel[np.absolute(kernel) < 0.01] = 0                                                  # zaf.py:551
    sparse = scipy.sparse.csr_matrix(kernel)                                                # zaf.py:554
    return np.conjugate(sparse) / length                                                    # zaf.py:557

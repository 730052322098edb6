"""NumPy models of the warp-level four-step FFTs in zaf-python_b200/csrc/fft_core.cuh, following the comments there
line by line (which element a lane holds, which twiddle it applies, where a result ends up), checked against
numpy.fft.  They pin the index maps the CUDA code implements -- the kernels themselves are checked on the GPU."""
import numpy as np
import pytest


def bitrev(v, bits):
    return int(format(v, f"0{bits}b")[::-1], 2) if bits else 0


def four_step(x, regs):
    """M = regs x 32 points, one 'warp': lane holds x[lane + 32 r].  Step 1: FFT over r; twiddle W_M^{k1 n2};
    step 2: FFT-32 over the lane index for every k1.  Result X[k1 + regs k2]."""
    m = regs * 32
    v = x.reshape(regs, 32)                                  # v[r, lane]
    y = np.fft.fft(v, axis=0)                                # y[k1, n2]
    y = y * np.exp(-2j * np.pi * np.outer(np.arange(regs), np.arange(32)) / m)
    z = np.fft.fft(y, axis=1)                                # z[k1, k2] = X[k1 + regs k2]
    out = np.empty(m, complex)
    for k1 in range(regs):
        out[k1 + regs * np.arange(32)] = z[k1]
    return out


def warp_fft128_model(x):
    """warp_fft128: step 2 is shared by lane octets (k1 = lane & 3, o = lane >> 2), n2 = a + 4 b, k2 = o + 8 j:
    X[k1 + 4 k2] = sum_a W_4^{a j} W_32^{a o} sum_b W_8^{b o} y[a + 4 b]; lane holds X[lane + 32 j] in v[bitrev(j, 2)]."""
    y = np.fft.fft(x.reshape(4, 32), axis=0) * np.exp(-2j * np.pi * np.outer(np.arange(4), np.arange(32)) / 128)
    regs = np.empty((32, 4), complex)                        # regs[lane, register]
    for lane in range(32):
        k1, o = lane & 3, lane >> 2
        u = np.array([sum(np.exp(-2j * np.pi * b * o / 8) * y[k1, a + 4 * b] for b in range(8)) for a in range(4)])
        u = u * np.exp(-2j * np.pi * np.arange(4) * o / 32)
        f = np.fft.fft(u)                                    # over a -> j
        for j in range(4):
            regs[lane, bitrev(j, 2)] = f[j]
    out = np.empty(128, complex)
    for lane in range(32):
        for j in range(4):
            out[lane + 32 * j] = regs[lane, bitrev(j, 2)]
    return out


def warp_fft2048_model(x):
    """warp_fft2048: 64 x 32; after the transpose lane l holds rows k1 = l (v[0..31]) and l + 32 (v[32..63]); outputs
    X[k1 + 64 k2] = X[lane + 32 (h + 2 k2)] land in v[32 h + bitrev(k2, 5)] == v[bitrev(h + 2 k2, 6)]."""
    y = np.fft.fft(x.reshape(64, 32), axis=0) * np.exp(-2j * np.pi * np.outer(np.arange(64), np.arange(32)) / 2048)
    regs = np.empty((32, 64), complex)
    for lane in range(32):
        for h in (0, 1):
            f = np.fft.fft(y[lane + 32 * h])
            for k2 in range(32):
                regs[lane, 32 * h + bitrev(k2, 5)] = f[k2]
    out = np.empty(2048, complex)
    for lane in range(32):
        for k in range(64):                                  # the kernels' convention: X[lane + 32 k] = v[bitrev(k, 6)]
            out[lane + 32 * k] = regs[lane, bitrev(k, 6)]
    return out


@pytest.mark.parametrize("regs", [4, 8, 16, 32, 64])
def test_four_step_decomposition(regs):
    rng = np.random.default_rng(regs)
    x = rng.standard_normal(regs * 32) + 1j * rng.standard_normal(regs * 32)
    assert np.allclose(four_step(x, regs), np.fft.fft(x), atol=1e-10)


def test_warp_fft128_and_2048_register_maps():
    rng = np.random.default_rng(1)
    for model, m in ((warp_fft128_model, 128), (warp_fft2048_model, 2048)):
        x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
        assert np.allclose(model(x), np.fft.fft(x), atol=1e-9), m


def test_real_input_unpack_and_mirror():
    """The N-point spectrum of a real frame from the N/2-point FFT of z[n] = x[2n] + i x[2n+1] (stft_warp_kernel):
    E = Z[k] + conj(Z[M-k]), O = -i (Z[k] - conj(Z[M-k])), X[k] = (E + W_N^k O) / 2, X[k+M] = (E - W_N^k O) / 2 for
    k < M/2, X[M/2] = conj-pair of Z[M/2], and the other half as the conjugate mirror X[N-k] = conj(X[k])."""
    rng = np.random.default_rng(2)
    for n in (256, 2048):
        m = n // 2
        x = rng.standard_normal(n)
        z = np.fft.fft(x[0::2] + 1j * x[1::2])
        got = np.empty(n, complex)
        for k in range(m // 2):
            p = np.conj(z[(m - k) % m])
            e, o = z[k] + p, -1j * (z[k] - p)
            t = np.exp(-2j * np.pi * k / n) * o
            got[k], got[k + m] = (e + t) / 2, (e - t) / 2
            if k > 0:
                got[m - k], got[n - k] = np.conj(got[k + m]), np.conj(got[k])
        got[m // 2] = np.conj(z[m // 2])
        got[m + m // 2] = z[m // 2]
        assert np.allclose(got, np.fft.fft(x), atol=1e-10), n

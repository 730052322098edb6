"""CPU-only checks of the boundary: the shared library loads, exports every symbol that
include/zafb200.h declares, the integer bookkeeping is bit-exact, the host-side operator
constructors reproduce the reference's operators, and compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def zaf():
    import __graft_entry__ as g

    g.build()
    import zaf_python_b200 as z

    return z


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "zafb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zafb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(zaf):
    names = _declared_functions()
    assert len(names) >= 40
    handle = ctypes.CDLL(zaf._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    # the ctypes prototype table mirrors the header one to one
    assert sorted(zaf._lib.PROTOTYPES) == names


def test_version_and_error_string(zaf):
    assert b"sm_100a" in zaf._lib.lib().zafb_version()
    with pytest.raises(ValueError):
        zaf.stft_geometry(10, 0, 1)
    assert b"stft geometry" in zaf._lib.lib().zafb_last_error()


def test_geometry_bit_exact_vs_reference(zaf, golden):
    g = golden("geometry")
    for ns, n, hop, nt, ilen in g.raw("stft_rows"):
        pad, nt_c, tail = zaf.stft_geometry(ns, n, hop)
        assert (pad, nt_c, tail) == oracle.stft_geometry(int(ns), int(n), int(hop))
        assert nt_c == nt
        assert zaf.istft_geometry(n, nt, hop)[2] == ilen
    for ns, n, nt, ilen in g.raw("mdct_rows"):
        half, nt_c, tail = zaf.mdct_geometry(ns, n)
        assert (half, nt_c) == (n // 2, nt) and tail == oracle.mdct_geometry(int(ns), int(n))[3]
        assert zaf.imdct_geometry(half, nt)[1] == ilen


def test_geometry_sweep_vs_oracle(zaf):
    rng = np.random.default_rng(7)
    for _ in range(2000):
        n = int(rng.integers(1, 5000))
        hop = int(rng.integers(1, n + 1))
        ns = int(rng.integers(0, 200000))
        assert zaf.stft_geometry(ns, n, hop) == oracle.stft_geometry(ns, n, hop)
        nt = oracle.stft_geometry(ns, n, hop)[1]
        total, trim, length = oracle.istft_length(n, nt, hop)
        assert zaf.istft_geometry(n, nt, hop) == (total, trim, length)
        if n >= 2:
            half, nt_m, _, tail = oracle.mdct_geometry(ns, n)
            assert zaf.mdct_geometry(ns, n) == (half, nt_m, tail)
            assert zaf.imdct_geometry(half, nt_m) == oracle.imdct_length(half, nt_m)
    for fs, tr, ns, L in ((44100, 25, 882000, 32768), (44100, 8, 100000, 32768), (16000, 100, 5000, 4096),
                          (22050, 7, 12345, 16384), (44100, 24, 10, 8192)):
        assert zaf.cqt_geometry(ns, fs, tr, L) == oracle.cqt_geometry(ns, fs, tr, L)
    # BASELINE configs
    assert zaf.stft_geometry(480000, 2048, 1024)[1] == 470
    assert zaf.stft_geometry(480000, 2048, 512)[1] == 939
    assert zaf.stft_geometry(80000, 1024, 256)[1] == 314
    assert zaf.mdct_geometry(1323000, 2048)[1] == 1293
    assert zaf.cqt_geometry(882000, 44100, 25, 32768)[:2] == (1764, 500)


def test_operator_constructors_match_reference(zaf, golden):
    g = golden("mel")
    assert np.array_equal(zaf.melfilterbank(16000, 1024, 128).toarray(), g.raw("fb_16k_1024_128"))
    assert np.array_equal(zaf.melfilterbank(44100, 2048, 128).toarray(), g.raw("fb_44k_2048_128"))
    assert np.array_equal(zaf.melfilterbank(8000, 256, 20).toarray(), g.raw("fb_8k_256_20"))
    gc = golden("cqt")
    for tag in gc.cases():
        fs, res, fmin, fmax = gc.get(tag, "params")
        k = zaf.cqtkernel(int(fs), int(res), fmin, fmax)
        assert k.shape == tuple(gc.get(tag, "kernel_shape"))
        assert np.array_equal(k.indptr, gc.get(tag, "kernel_indptr"))
        assert np.array_equal(k.indices, gc.get(tag, "kernel_indices"))
        assert np.allclose(k.data, gc.get(tag, "kernel_data"), rtol=0, atol=1e-18)


def test_operator_memo_reuses_the_prepared_operator_per_object(zaf):
    """melspectrogram / cqtspectrogram calls in a loop pass the same operator object: its densified / sorted form and
    its plan-cache key are prepared once per object, again after an in-place edit, never for a different object."""
    memo, calls = zaf._OperatorMemo(), []

    def prep(o):
        calls.append(1)
        return zaf._prepare_filterbank(o)

    fb = zaf.melfilterbank(16000, 1024, 128)
    first = memo.get(fb, prep)
    assert memo.get(fb, prep) is first and len(calls) == 1
    assert np.array_equal(first[0], fb.toarray()) and first[1] == first[0].tobytes()
    fb.data[0] += 1.0  # edited in place: the fingerprint changes
    assert memo.get(fb, prep) is not first and len(calls) == 2
    other = zaf.melfilterbank(16000, 1024, 128)
    assert np.array_equal(memo.get(other, prep)[0], other.toarray()) and len(calls) == 3
    memo.get([[1.0, 2.0]], prep)  # objects without weak references are prepared every time
    memo.get([[1.0, 2.0]], prep)
    assert len(calls) == 5
    kern = zaf.cqtkernel(44100, 12, 32.70, 4186.01)
    shape, data, indptr, indices, key = zaf._cqt_ops.get(kern, zaf._prepare_cqt_kernel)
    assert zaf._cqt_ops.get(kern, zaf._prepare_cqt_kernel)[1] is data
    assert shape == kern.shape and len(data) == kern.nnz and indptr[-1] == kern.nnz and len(key) == 3


def test_no_cpu_fallback(zaf):
    """Without a CUDA device every compute entry point raises; nothing is computed on the host."""
    if zaf.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        zaf.stft(np.zeros(1000, np.float32), np.ones(64), 16)
    with pytest.raises(RuntimeError):
        zaf.mdct(np.zeros(1000, np.float32), np.ones(64))


def test_host_mirror_fill_is_the_hermitian_half(zaf):
    """zafb_host_mirror_fill (the host half of the half-spectrum D2H path, no device involved): bins N/2+1 .. N-1 become
    conj of bins N/2-1 .. 1, bit for bit, for aligned and unaligned buffers; bins 0 .. N/2 are left alone."""
    import ctypes as C
    lib = zaf._lib.lib()
    rng = np.random.default_rng(3)
    for n, frames, off in ((4, 3, 0), (8, 5, 0), (64, 7, 0), (2048, 3, 0), (1024, 4, 1), (12, 2, 1)):
        buf = np.zeros(frames * n + 1, np.complex64)
        a = buf[off:off + frames * n].reshape(frames, n)
        a[:] = (rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))).astype(np.complex64)
        want = a.copy()
        want[:, n // 2 + 1:] = np.conj(want[:, 1:n // 2][:, ::-1])
        zaf._lib.check(lib.zafb_host_mirror_fill(C.c_void_p(a.ctypes.data), frames, n))
        assert np.array_equal(a.view(np.uint32), want.view(np.uint32)), (n, frames, off)
    with pytest.raises(ValueError):
        zaf._lib.check(lib.zafb_host_mirror_fill(C.c_void_p(buf.ctypes.data), 1, 6))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "zaf-python_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f

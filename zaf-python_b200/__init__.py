"""zaf-python_b200: B200 (sm_100a) drop-in for the transform hot path of zaf.py.

    import zaf_python_b200 as zaf          # repo-root shim for the hyphenated directory name
    X = zaf.stft(x, w, hop)                # same signatures as the reference module

Every function keeps the positional signature of its reference counterpart (zaf.py line cited in
each docstring) and returns arrays of the same *shape*; values are computed in fp32 on the GPU
(complex64 / float32 results) by the hand-written CUDA kernels in ``csrc/`` through the C ABI in
``include/zafb200.h``.  Extensions, all keyword-only or by type:

* a leading batch axis: ``(B, number_samples)`` in -> ``(B, ...)`` out (the reference rejects 2-D);
* ``DeviceArray`` in -> ``DeviceArray`` out (device-resident pipelines, no host round trip);
* ``layout="frame_major"`` (default: memory is ``[frame][bin]`` and the result is the transposed
  view, i.e. the reference's shape with Fortran-like strides) or ``layout="bin_major"`` (the
  reference's C-order memory).

There is no CPU fallback and no dependency on PyTorch/CuPy/cuFFT: without the shared library or a
CUDA device the functions raise.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict

import numpy as np

from . import _lib, _pinned
from ._device import (DeviceArray, Event, PinnedArray, Stream, device_count, empty, ensure_init, init,  # noqa: F401
                      host_copy_bytes, launch_count, synchronize, to_device)
from ._lib import LAYOUT_BIN_MAJOR, LAYOUT_FRAME_MAJOR, ZafbError  # noqa: F401
from ._operators import cqtkernel, melfilterbank  # noqa: F401
from . import _dist as dist  # noqa: F401
from ._dist import shard_range  # noqa: F401

__all__ = [
    "stft", "istft", "melfilterbank", "melspectrogram", "mfcc", "cqtkernel", "cqtspectrogram",
    "cqtchromagram", "dct", "dst", "mdct", "imdct", "init", "device_count", "synchronize",
    "to_device", "empty", "from_pcm16", "DeviceArray", "PinnedArray", "Stream", "Event", "launch_count", "host_copy_bytes",
    "spec_abs", "spec_mask", "spec_mirror", "ratio_min", "multiply", "quantize", "count_mismatch",
    "stft_geometry", "istft_geometry", "mdct_geometry", "imdct_geometry", "cqt_geometry", "dist", "shard_range",
]

_LAYOUTS = {"frame_major": LAYOUT_FRAME_MAJOR, "bin_major": LAYOUT_BIN_MAJOR}


# ------------------------------------------------------------------ integer bookkeeping
def _geom(fn, n_out, *args):
    outs = [C.c_int64(0) for _ in range(n_out)]
    _lib.check(getattr(_lib.lib(), fn)(*[int(a) for a in args], *[C.byref(o) for o in outs]))
    return tuple(o.value for o in outs)


def stft_geometry(number_samples, window_length, step_length):
    """(front pad, number_times, tail pad) -- zaf.py:99-121, bit-exact."""
    return _geom("zafb_stft_geometry", 3, number_samples, window_length, step_length)


def istft_geometry(window_length, number_times, step_length):
    """(overlap-add length, slice start, number_samples) -- zaf.py:217, 236-238."""
    return _geom("zafb_istft_geometry", 3, window_length, number_times, step_length)


def mdct_geometry(number_samples, window_length):
    """(M, number_times, tail pad) -- zaf.py:1029-1041."""
    return _geom("zafb_mdct_geometry", 3, number_samples, window_length)


def imdct_geometry(number_frequencies, number_times):
    """(overlap-add length, number_samples) -- zaf.py:1132, 1182."""
    return _geom("zafb_imdct_geometry", 2, number_frequencies, number_times)


def cqt_geometry(number_samples, sampling_frequency, time_resolution, fft_length):
    """(step, number_times, front pad, back pad) -- zaf.py:603-620.  ``round`` is Python's
    round-half-to-even on the float quotient, as in the reference."""
    step = round(sampling_frequency / time_resolution)
    return (step,) + _geom("zafb_cqt_geometry", 3, number_samples, step, fft_length)


# ------------------------------------------------------------------ PCM input
def from_pcm16(pcm, *, mono=False, stream=None):
    """int16 PCM, shape (number_samples,) or (number_samples, number_channels) as ``scipy.io.wavfile.read`` returns it,
    to a float32 DeviceArray normalised like ``zaf.wavread`` (zaf.py:1199-1202: x / 2**15), converted ON the GPU so only
    the 16-bit samples cross PCIe.  Result: (number_channels, number_samples) -- one clip per channel, ready for the
    batched transforms -- or, with ``mono=True``, the channel mean (number_samples,) of the reference's examples."""
    a = np.ascontiguousarray(pcm)
    if a.dtype != np.int16 or a.ndim not in (1, 2):
        raise ValueError("pcm must be an int16 array of shape (number_samples,) or (number_samples, number_channels)")
    frames = a.shape[0]
    channels = 1 if a.ndim == 1 else a.shape[1]
    ensure_init()
    raw = to_device(a.reshape(-1), stream=stream)
    pitch = _even(frames)
    if mono or a.ndim == 1:
        out = DeviceArray((frames,), np.float32)
    else:
        out = DeviceArray((channels, pitch), np.float32, cols=frames)
    try:
        _lib.check(_lib.lib().zafb_pcm16_to_f32(C.c_void_p(raw.ptr), frames, channels, 1 if (mono or a.ndim == 1) else 0,
                                                C.c_void_p(out.ptr), pitch, _stream_ptr(stream)))
        if stream is None:
            synchronize()
        else:
            stream.synchronize()
    finally:
        raw.free()
    return out


# ------------------------------------------------------------------ plan caches
class _PlanCache:
    def __init__(self, create, destroy, capacity=16):
        self._create, self._destroy, self._cap = create, destroy, capacity
        self._d = OrderedDict()

    def get(self, key, *args):
        p = self._d.get(key)
        if p is not None:
            self._d.move_to_end(key)
            return p
        ensure_init()
        handle = C.c_void_p()
        _lib.check(getattr(_lib.lib(), self._create)(C.byref(handle), *args))
        self._d[key] = handle
        while len(self._d) > self._cap:
            _, old = self._d.popitem(last=False)
            getattr(_lib.lib(), self._destroy)(old)
        return handle


_stft_plans = _PlanCache("zafb_stft_plan_create", "zafb_stft_plan_destroy")
_mdct_plans = _PlanCache("zafb_mdct_plan_create", "zafb_mdct_plan_destroy")
_dct_plans = _PlanCache("zafb_dct_plan_create", "zafb_dct_plan_destroy", capacity=64)
_mel_plans = _PlanCache("zafb_mel_plan_create", "zafb_mel_plan_destroy", capacity=8)
_cqt_plans = _PlanCache("zafb_cqt_plan_create", "zafb_cqt_plan_destroy", capacity=4)


def _window64(window_function):
    w = np.ascontiguousarray(window_function, dtype=np.float64)
    if w.ndim != 1:
        raise ValueError("window_function must be 1-D")
    return w


def _stft_plan(window_function, step_length):
    w = _window64(window_function)
    hop = int(step_length)
    return _stft_plans.get(("stft", len(w), hop, w.tobytes()), w.ctypes.data, len(w), hop), w


def _layout_id(layout):
    try:
        return _LAYOUTS[layout]
    except KeyError:
        raise ValueError(f"layout must be one of {sorted(_LAYOUTS)}") from None


def _signal_batch(audio_signal):
    """-> (float32 C-contiguous (B, ns) array, was_1d).  Like the reference, >2-D is an error."""
    x = np.asarray(audio_signal)
    if x.ndim not in (1, 2):
        raise ValueError("audio_signal must have shape (number_samples,) or (batch, number_samples)")
    one = x.ndim == 1
    x = np.ascontiguousarray(x.reshape(1, -1) if one else x, dtype=np.float32)
    return x, one


def _matrix_out(batch, rows, cols, dtype, layout, one):
    """Host result buffer for a (rows, cols) = (bins, frames) matrix per clip, and the view to return."""
    if layout == LAYOUT_FRAME_MAJOR:
        mem = _pinned.empty((batch, cols, rows), dtype)
        view = np.swapaxes(mem, 1, 2)
    else:
        mem = _pinned.empty((batch, rows, cols), dtype)
        view = mem
    return mem, (view[0] if one else view)


def _even(n):
    return (int(n) + 1) & ~1


def _stream_ptr(stream):
    return stream.ptr if stream is not None else None


def _device_signal(x):
    """Validate a device-resident signal batch: float32, (ns,) or (B, ns), not a transposed view."""
    if x.dtype != np.float32 or len(x.shape) not in (1, 2) or x.transposed:
        raise ValueError("device input must be a float32 (number_samples,) or (batch, number_samples) DeviceArray")
    one = len(x.shape) == 1
    batch, ns = (1, x.shape[0]) if one else x.shape
    return one, batch, ns


def _device_matrix(s, dtype, what):
    """Validate a device-resident (bins, frames) / (B, bins, frames) spectrum of the given dtype."""
    if s.dtype != np.dtype(dtype) or len(s.shape) not in (2, 3) or s.cols is not None:
        raise ValueError(f"device input must be a {np.dtype(dtype)} {what} DeviceArray")
    one = len(s.shape) == 2
    return one, (1 if one else s.shape[0])


# ------------------------------------------------------------------ STFT / ISTFT
def _stft_onesided(plan, w, audio_signal, step_length, layout, stream):
    """Bins 0 .. N/2 only (non-reference extension), frame-major memory [clip][frame][N/2+1]."""
    if layout != "frame_major":
        raise ValueError("onesided=True returns frame-major memory (layout='frame_major')")
    n = len(w)
    bins = n // 2 + 1
    on_device = isinstance(audio_signal, DeviceArray)
    if on_device:
        x = audio_signal
        one, batch, ns = _device_signal(x)
    else:
        host, one = _signal_batch(audio_signal)
        batch, ns = host.shape
        x = to_device(host, stream=stream)
    nt = stft_geometry(ns, n, step_length)[1]
    out = DeviceArray((nt, bins) if one else (batch, nt, bins), np.complex64, transposed=True)
    _lib.check(_lib.lib().zafb_stft_onesided_f32(plan, C.c_void_p(x.ptr), batch, ns, x.pitch, C.c_void_p(out.ptr), bins,
                                                 _stream_ptr(stream)))
    if on_device:
        return out
    res = out.to_host(stream=stream)
    if stream is not None:
        stream.synchronize()
    x.free()
    out.free()
    return res


def stft(audio_signal, window_function, step_length, *, layout="frame_major", stream=None, out=None, onesided=False):
    """Short-time Fourier transform -- drop-in for ``zaf.stft`` (zaf.py:45-141).

    Returns the full two-sided spectrum of shape (window_length, number_times) [complex64];
    centre padding floor(N/2), number of frames and tail padding follow zaf.py:99-121 exactly.
    ``out=`` supplies the result memory: a NumPy array (e.g. pinned) on the host path, a DeviceArray
    on the device path.  ``onesided=True`` is an explicit NON-reference mode: only rows 0 .. N/2 are
    computed and returned, shape (window_length/2 + 1, number_times) -- the other rows of the reference's
    result are their conjugate mirror (``spec_mirror`` rebuilds them) -- at half the spectrum's memory traffic.
    """
    plan, w = _stft_plan(window_function, step_length)
    if onesided:
        if out is not None:
            raise ValueError("out= is not supported with onesided=True")
        return _stft_onesided(plan, w, audio_signal, step_length, layout, stream)
    lay = _layout_id(layout)
    n = len(w)
    if isinstance(audio_signal, DeviceArray):
        x = audio_signal
        one, batch, ns = _device_signal(x)
        nt = stft_geometry(ns, n, step_length)[1]
        mem_shape = (batch, nt, n) if lay == LAYOUT_FRAME_MAJOR else (batch, n, nt)
        if out is None:
            out = DeviceArray(mem_shape[1:] if one else mem_shape, np.complex64, transposed=lay == LAYOUT_FRAME_MAJOR)
        elif not isinstance(out, DeviceArray) or out.dtype != np.complex64 or out.nbytes != 8 * batch * nt * n:
            raise ValueError(f"out must be a complex64 DeviceArray with {batch * nt * n} elements")
        else:  # caller-provided device memory, viewed in the requested layout
            out = DeviceArray(mem_shape[1:] if one else mem_shape, np.complex64, ptr=out.ptr, owner=out,
                              transposed=lay == LAYOUT_FRAME_MAJOR)
        _lib.check(_lib.lib().zafb_stft_f32(plan, C.c_void_p(x.ptr), batch, ns, x.pitch, C.c_void_p(out.ptr), lay,
                                            _stream_ptr(stream)))
        return out
    x, one = _signal_batch(audio_signal)
    batch, ns = x.shape
    nt = stft_geometry(ns, n, step_length)[1]
    if out is None:
        mem, view = _matrix_out(batch, n, nt, np.complex64, lay, one)
    else:  # caller-provided (e.g. pinned) result memory in the chosen layout
        want = (batch, nt, n) if lay == LAYOUT_FRAME_MAJOR else (batch, n, nt)
        if out.dtype != np.complex64 or not out.flags.c_contiguous or out.size != int(np.prod(want)):
            raise ValueError(f"out must be a C-contiguous complex64 array with {want} elements")
        mem = out.reshape(want)
        view = np.swapaxes(mem, 1, 2) if lay == LAYOUT_FRAME_MAJOR else mem
        view = view[0] if one else view
    _lib.check(_lib.lib().zafb_stft_host_f32(plan, x.ctypes.data, batch, ns, ns, mem.ctypes.data, lay))
    return view


def _spec_memory(audio_stft, dtype):
    """Find the memory layout of a (..., bins, frames) host array without copying if possible."""
    a = np.asarray(audio_stft)
    if a.ndim not in (2, 3):
        raise ValueError("expected shape (bins, frames) or (batch, bins, frames)")
    one = a.ndim == 2
    if one:
        a = a[None]
    if a.dtype != dtype:
        a = a.astype(dtype)
    t = np.swapaxes(a, 1, 2)
    if t.flags.c_contiguous and not a.flags.c_contiguous:
        return t, LAYOUT_FRAME_MAJOR, one, a.shape
    return np.ascontiguousarray(a), LAYOUT_BIN_MAJOR, one, a.shape


def _istft_onesided(plan, w, audio_stft, step_length, stream, mask=None):
    n = len(w)
    bins = n // 2 + 1
    on_device = isinstance(audio_stft, DeviceArray)
    if mask is not None and not on_device:
        raise ValueError("mask= needs DeviceArray spectrum and mask (device-resident chains)")
    if on_device:
        s = audio_stft
        one, batch = _device_matrix(s, np.complex64, "(N/2+1, nt) or (B, N/2+1, nt)")
        if not s.transposed:
            raise ValueError("a one-sided device spectrum must be frame-major (as stft(..., onesided=True) returns it)")
        shape = s.shape
    else:
        a = np.asarray(audio_stft)
        if a.ndim not in (2, 3):
            raise ValueError("expected shape (N/2+1, nt) or (batch, N/2+1, nt)")
        one = a.ndim == 2
        a = a[None] if one else a
        batch, shape = a.shape[0], a.shape
        mem = np.ascontiguousarray(np.swapaxes(a, 1, 2), dtype=np.complex64)  # frame-major memory
        d = to_device(mem, stream=stream)
        s = DeviceArray(d.mem_shape, np.complex64, ptr=d.ptr, owner=d, transposed=True)
    if shape[-2] != bins:
        raise ValueError(f"a one-sided spectrum has {bins} rows for a window of {n} samples, not {shape[-2]}")
    nt = shape[-1]
    length = istft_geometry(n, nt, step_length)[2]
    pitch = _even(length)
    out = DeviceArray((pitch,) if one else (batch, pitch), np.float32, cols=length)
    if mask is not None:
        mp = _mask_pitch(mask, s, n)
        try:
            _lib.check(_lib.lib().zafb_istft_masked_f32(plan, C.c_void_p(s.ptr), batch, nt, bins, 1, C.c_void_p(mask.ptr), mp,
                                                        C.c_void_p(out.ptr), pitch, _stream_ptr(stream)))
            return out
        except NotImplementedError:  # no fused kernel for this geometry: multiply the half spectrum, then transform
            s = spec_mask(s, mask, stream=stream)
    _lib.check(_lib.lib().zafb_istft_onesided_f32(plan, C.c_void_p(s.ptr), batch, nt, bins, C.c_void_p(out.ptr), pitch,
                                                  _stream_ptr(stream)))
    if on_device:
        return out
    y = out.to_host(stream=stream)
    if stream is not None:
        stream.synchronize()
    out.free()
    return y


def spec_mirror(audio_stft_onesided, window_length, *, stream=None):
    """The reference's two-sided spectrum (window_length rows) from a one-sided, frame-major device spectrum:
    rows N-k = conj(rows k), ON the device."""
    s = audio_stft_onesided
    n = int(window_length)
    one, batch = _device_matrix(s, np.complex64, "(N/2+1, nt) or (B, N/2+1, nt)")
    if not s.transposed or s.shape[-2] != n // 2 + 1:
        raise ValueError("expected the frame-major (N/2+1, nt) result of stft(..., onesided=True)")
    nt = s.shape[-1]
    out = DeviceArray((nt, n) if one else (batch, nt, n), np.complex64, transposed=True)
    _lib.check(_lib.lib().zafb_spec_mirror_f32(C.c_void_p(s.ptr), n // 2 + 1, batch * nt, n, C.c_void_p(out.ptr), _stream_ptr(stream)))
    return out


def _mask_pitch(mask, spec, n):
    """Validate a (N/2+1, nt) / (B, N/2+1, nt) float32 DeviceArray mask for ``spec``.  Returns its row pitch when both are
    frame-major (the layout the fused kernel reads), else None: the caller multiplies first (``spec_mask``)."""
    if not isinstance(mask, DeviceArray) or not isinstance(spec, DeviceArray):
        raise ValueError("mask= needs DeviceArray spectrum and mask (device-resident chains)")
    if mask.dtype != np.float32 or mask.transposed != spec.transposed:
        raise ValueError("mask= needs a float32 mask in the layout of the spectrum (as zaf.spec_abs / zaf.ratio_min return it)")
    if mask.shape[-2] != n // 2 + 1 or mask.shape[-1] != spec.shape[-1] or mask.shape[:-2] != spec.shape[:-2]:
        raise ValueError(f"mask must have {n // 2 + 1} rows and the spectrum's frames and batch, not {mask.shape}")
    return mask.mem_shape[-1] if mask.transposed else None


def istft(audio_stft, window_function, step_length, *, stream=None, onesided=False, mask=None):
    """Inverse STFT by constant overlap-add -- drop-in for ``zaf.istft`` (zaf.py:144-243).

    Output length nt*hop - (N - hop); only the real part of the inverse transform is kept, no
    synthesis window, division by sum(w[0:N:hop]) -- all as in the reference (including its
    N-hop trim, which makes the round trip an identity only for hop = N/2).  ``onesided=True``
    (non-reference mode) takes rows 0 .. N/2 only and treats the rest as their conjugate mirror.
    ``mask=`` (non-reference, DeviceArrays only): the transform of ``np.concatenate((mask, mask[-2:0:-1])) * X``
    (zaf.py:185-190) for a real mask with rows 0 .. N/2 -- fused into the ISTFT's loads for N = 2048, hop = 512, a
    ``spec_mask`` pass followed by the plain transform otherwise.
    """
    plan, w = _stft_plan(window_function, step_length)
    if onesided:
        return _istft_onesided(plan, w, audio_stft, step_length, stream, mask)
    n = len(w)
    if isinstance(audio_stft, DeviceArray):
        s = audio_stft
        one, batch = _device_matrix(s, np.complex64, "(N, nt) or (B, N, nt)")
        shape = s.shape
        if shape[-2] != n:
            raise ValueError(f"audio_stft has {shape[-2]} bins but the window has {n} samples")
        nt = shape[-1]
        length = istft_geometry(n, nt, step_length)[2]
        out = DeviceArray((length,) if one else (batch, length), np.float32)
        if mask is not None:
            mp = _mask_pitch(mask, s, n)
            try:
                if mp is None:
                    raise NotImplementedError
                _lib.check(_lib.lib().zafb_istft_masked_f32(plan, C.c_void_p(s.ptr), batch, nt, n, 0, C.c_void_p(mask.ptr), mp,
                                                            C.c_void_p(out.ptr), length, _stream_ptr(stream)))
                return out
            except NotImplementedError:  # no fused kernel for this geometry / layout: multiply, then transform
                s = spec_mask(s, mask, stream=stream)
        lay = LAYOUT_FRAME_MAJOR if s.transposed else LAYOUT_BIN_MAJOR
        _lib.check(_lib.lib().zafb_istft_f32(plan, C.c_void_p(s.ptr), batch, nt, lay, C.c_void_p(out.ptr), length,
                                             _stream_ptr(stream)))
        return out
    if mask is not None:
        raise ValueError("mask= needs DeviceArray spectrum and mask (device-resident chains)")
    mem, lay, one, shape = _spec_memory(audio_stft, np.complex64)
    batch, bins, nt = shape
    if bins != n:
        raise ValueError(f"audio_stft has {bins} bins but the window has {n} samples")
    length = istft_geometry(n, nt, step_length)[2]
    y = _pinned.empty((batch, length), np.float32)
    _lib.check(_lib.lib().zafb_istft_host_f32(plan, mem.ctypes.data, batch, nt, lay, y.ctypes.data, length))
    return y[0] if one else y


# ------------------------------------------------------------------ device-resident stages between transforms
def _spec_geometry(spec):
    """(n_clips, bins, frames, layout id) of a complex64 / float32 (bins, frames) or (B, bins, frames) DeviceArray."""
    if not isinstance(spec, DeviceArray) or len(spec.shape) not in (2, 3) or spec.cols is not None:
        raise ValueError("expected a (bins, frames) or (batch, bins, frames) DeviceArray")
    shape = spec.shape
    clips = 1 if len(shape) == 2 else shape[0]
    return clips, shape[-2], shape[-1], (LAYOUT_FRAME_MAJOR if spec.transposed else LAYOUT_BIN_MAJOR)


def _like(spec, bins, dtype):
    """A new DeviceArray with `bins` rows in the layout (and batch shape) of ``spec``."""
    lead = spec.mem_shape[:-2]
    frames = spec.shape[-1]
    mem = lead + ((frames, bins) if spec.transposed else (bins, frames))
    return DeviceArray(mem, dtype, transposed=spec.transposed)


def spec_abs(audio_stft, number_frequencies=None, *, stream=None):
    """|X| of rows 0 .. number_frequencies-1 of a device-resident spectrum, ON the device -- the
    ``abs(audio_stft[0:number_frequencies, :])`` of the reference's examples (zaf.py:176-177).  Same layout as the input."""
    clips, bins, frames, lay = _spec_geometry(audio_stft)
    if audio_stft.dtype != np.complex64:
        raise ValueError("spec_abs needs a complex64 spectrum")
    keep = bins if number_frequencies is None else int(number_frequencies)
    out = _like(audio_stft, keep, np.float32)
    _lib.check(_lib.lib().zafb_spec_abs_f32(C.c_void_p(audio_stft.ptr), clips, bins, frames, lay, keep, C.c_void_p(out.ptr),
                                            _stream_ptr(stream)))
    return out


def spec_mask(audio_stft, mask, *, out=None, stream=None):
    """``audio_stft * mask`` ON the device.  ``mask`` (float32 DeviceArray, same layout) has either every row of the
    spectrum or rows 0 .. N/2, in which case it is mirrored onto the upper half exactly like the reference's
    ``np.concatenate((mask, mask[-2:0:-1, :]))`` (zaf.py:185-186).  ``out=audio_stft`` works in place."""
    clips, bins, frames, lay = _spec_geometry(audio_stft)
    mclips, mbins, mframes, mlay = _spec_geometry(mask)
    if audio_stft.dtype != np.complex64 or mask.dtype != np.float32:
        raise ValueError("spec_mask needs a complex64 spectrum and a float32 mask")
    if (mclips, mframes, mlay) != (clips, frames, lay):
        raise ValueError("mask and spectrum must agree in batch size, number of frames and layout")
    if out is None:
        out = _like(audio_stft, bins, np.complex64)
    elif out.dtype != np.complex64 or out.mem_shape != audio_stft.mem_shape or out.transposed != audio_stft.transposed:
        raise ValueError("out must match the spectrum")
    _lib.check(_lib.lib().zafb_spec_mask_f32(C.c_void_p(audio_stft.ptr), clips, bins, frames, lay, C.c_void_p(mask.ptr), mbins,
                                             C.c_void_p(out.ptr), _stream_ptr(stream)))
    return out


def _same_real(a, b):
    if not (isinstance(a, DeviceArray) and isinstance(b, DeviceArray)) or a.dtype != np.float32 or b.dtype != np.float32 \
            or a.mem_shape != b.mem_shape or a.transposed != b.transposed:
        raise ValueError("operands must be float32 DeviceArrays of the same shape and layout")


def _real_like(a):
    return DeviceArray(a.mem_shape, np.float32, transposed=a.transposed, cols=a.cols)


def ratio_min(a, b, *, stream=None):
    """``np.minimum(a, b) / a`` ON the device -- the centre-mask estimate of zaf.py:181-182."""
    _same_real(a, b)
    out = _real_like(a)
    _lib.check(_lib.lib().zafb_ratio_min_f32(C.c_void_p(a.ptr), C.c_void_p(b.ptr), a.nbytes // 4, C.c_void_p(out.ptr),
                                             _stream_ptr(stream)))
    return out


def multiply(a, b, *, out=None, stream=None):
    """``a * b`` for float32 DeviceArrays (e.g. a mask on MDCT coefficients), ON the device."""
    _same_real(a, b)
    out = _real_like(a) if out is None else out
    _lib.check(_lib.lib().zafb_mul_f32(C.c_void_p(a.ptr), C.c_void_p(b.ptr), a.nbytes // 4, C.c_void_p(out.ptr),
                                       _stream_ptr(stream)))
    return out


def quantize(x, step, *, out=None, stream=None):
    """Uniform scalar quantise-dequantise ``step * np.round(x / step)`` (float32, round half to even) ON the device: the
    coefficient-domain stage of an ``mdct -> ... -> imdct`` codec chain (zaf.py:1098-1105 is the chain without it)."""
    if not isinstance(x, DeviceArray) or x.dtype != np.float32:
        raise ValueError("quantize needs a float32 DeviceArray")
    out = _real_like(x) if out is None else out
    _lib.check(_lib.lib().zafb_quantize_f32(C.c_void_p(x.ptr), x.nbytes // 4, float(step), C.c_void_p(out.ptr),
                                            _stream_ptr(stream)))
    return out


def count_mismatch(a, b, *, stream=None):
    """Number of differing 32-bit words between two device buffers of equal size (bitwise comparison on the device)."""
    if a.nbytes != b.nbytes:
        raise ValueError("buffers differ in size")
    n = C.c_int64(0)
    _lib.check(_lib.lib().zafb_count_mismatch_u32(C.c_void_p(a.ptr), C.c_void_p(b.ptr), a.nbytes // 4, C.byref(n),
                                                  _stream_ptr(stream)))
    return int(n.value)


# ------------------------------------------------------------------ MDCT / IMDCT
def _mdct_plan(window_function):
    w = _window64(window_function)
    if len(w) % 2:
        raise ValueError("window_function must have an even length (zaf.py:1071 fails to broadcast otherwise)")
    return _mdct_plans.get(("mdct", len(w), w.tobytes()), w.ctypes.data, len(w)), w


def mdct(audio_signal, window_function, *, layout="frame_major", stream=None):
    """Modified discrete cosine transform -- drop-in for ``zaf.mdct`` (zaf.py:984-1075).
    Returns (window_length/2, number_times) float32."""
    plan, w = _mdct_plan(window_function)
    lay = _layout_id(layout)
    n = len(w)
    if isinstance(audio_signal, DeviceArray):
        x = audio_signal
        one, batch, ns = _device_signal(x)
        m, nt, _ = mdct_geometry(ns, n)
        mem_shape = (batch, nt, m) if lay == LAYOUT_FRAME_MAJOR else (batch, m, nt)
        out = DeviceArray(mem_shape[1:] if one else mem_shape, np.float32, transposed=lay == LAYOUT_FRAME_MAJOR)
        _lib.check(_lib.lib().zafb_mdct_f32(plan, C.c_void_p(x.ptr), batch, ns, x.pitch, C.c_void_p(out.ptr), lay,
                                            _stream_ptr(stream)))
        return out
    x, one = _signal_batch(audio_signal)
    batch, ns = x.shape
    m, nt, _ = mdct_geometry(ns, n)
    mem, view = _matrix_out(batch, m, nt, np.float32, lay, one)
    ensure_init()
    _lib.check(_lib.lib().zafb_mdct_host_f32(plan, x.ctypes.data, batch, ns, ns, mem.ctypes.data, lay))
    return view


def imdct(audio_mdct, window_function, *, stream=None):
    """Inverse MDCT with TDAC overlap-add -- drop-in for ``zaf.imdct`` (zaf.py:1078-1184).
    Output length M*(nt-1)-1 (the reference's ``[M : -M-1]`` slice)."""
    plan, w = _mdct_plan(window_function)
    n = len(w)
    if isinstance(audio_mdct, DeviceArray):
        s = audio_mdct
        one, batch = _device_matrix(s, np.float32, "(M, nt) or (B, M, nt)")
        shape = s.shape
        if shape[-2] * 2 != n:
            raise ValueError("audio_mdct rows must equal window_length/2")
        nt = shape[-1]
        length = imdct_geometry(n // 2, nt)[1]
        pitch = _even(length)  # the reference's odd length M(nt-1)-1 would misalign every other row
        out = DeviceArray((pitch,) if one else (batch, pitch), np.float32, cols=length)
        lay = LAYOUT_FRAME_MAJOR if s.transposed else LAYOUT_BIN_MAJOR
        _lib.check(_lib.lib().zafb_imdct_f32(plan, C.c_void_p(s.ptr), batch, nt, lay, C.c_void_p(out.ptr), pitch,
                                             _stream_ptr(stream)))
        return out
    mem, lay, one, shape = _spec_memory(audio_mdct, np.float32)
    batch, bins, nt = shape
    if bins * 2 != n:
        raise ValueError("audio_mdct rows must equal window_length/2")
    length = imdct_geometry(bins, nt)[1]
    y = _pinned.empty((batch, length), np.float32)
    _lib.check(_lib.lib().zafb_imdct_host_f32(plan, mem.ctypes.data, batch, nt, lay, y.ctypes.data, length))
    return y[0] if one else y


# ------------------------------------------------------------------ DCT / DST
def _dct_like(audio_signal, kind, dtype_code):
    if dtype_code not in (1, 2, 3, 4):
        return None  # the reference falls through its if/elif chain and returns None (zaf.py:759-839)
    if isinstance(audio_signal, DeviceArray):
        x = audio_signal
        one, batch, n = _device_signal(x)
        plan = _dct_plans.get((kind, dtype_code, n), kind, dtype_code, n)
        out = DeviceArray(x.shape, np.float32)
        _lib.check(_lib.lib().zafb_dct_f32(plan, C.c_void_p(x.ptr), batch, x.pitch, C.c_void_p(out.ptr), n, None))
        return out
    x, one = _signal_batch(audio_signal)
    batch, n = x.shape
    plan = _dct_plans.get((kind, dtype_code, n), kind, dtype_code, n)
    out = _pinned.empty((batch, n), np.float32)
    _lib.check(_lib.lib().zafb_dct_host_f32(plan, x.ctypes.data, batch, n, out.ctypes.data, n))
    return out[0] if one else out


def dct(audio_signal, dct_type):
    """Orthonormal DCT-I/II/III/IV -- drop-in for ``zaf.dct`` (zaf.py:703-839).  A 2-D input is a
    batch of vectors (extension).  Unknown types return None like the reference."""
    return _dct_like(audio_signal, 0, dct_type)


def dst(audio_signal, dst_type):
    """Orthonormal DST-I/II/III/IV -- drop-in for ``zaf.dst`` (zaf.py:842-981)."""
    return _dct_like(audio_signal, 1, dst_type)


# ------------------------------------------------------------------ mel spectrogram / MFCC
_MEL_ROUTES = {"fused": 0, "tensor": 1}


_MEL_PRECISIONS = {"auto": 0, "float32": 32, "float64": 64}


class _OperatorMemo:
    """Prepared form of an operator argument (densified float64 filterbank, sorted CSR arrays of a CQT kernel) + the
    bytes that key the plan cache, remembered per OBJECT: a caller that passes the same filterbank / kernel object again
    (the usual loop over clips) does not pay toarray() + tobytes() + hashing of the whole operator on every call.
    An entry is used only while the object is alive (weak reference) and its cheap fingerprint (shape, dtype, sum of the
    stored values) is unchanged, so an operator edited in place is prepared afresh."""

    def __init__(self, capacity=8):
        self._d = OrderedDict()
        self._cap = capacity

    @staticmethod
    def _fingerprint(obj):
        vals = obj.data if hasattr(obj, "indptr") else obj
        vals = np.asarray(vals)
        return (getattr(obj, "shape", None), str(vals.dtype), getattr(obj, "nnz", vals.size), complex(vals.sum()))

    def get(self, obj, prepare):
        try:
            ref = weakref.ref(obj)
        except TypeError:  # lists etc.: no memo
            return prepare(obj)
        hit = self._d.get(id(obj))
        fp = self._fingerprint(obj)
        if hit is not None and hit[0]() is obj and hit[1] == fp:
            self._d.move_to_end(id(obj))
            return hit[2]
        prepared = prepare(obj)
        self._d[id(obj)] = (ref, fp, prepared)
        while len(self._d) > self._cap:
            self._d.popitem(last=False)
        return prepared


_mel_ops = _OperatorMemo()
_cqt_ops = _OperatorMemo()


def _prepare_filterbank(mel_filterbank):
    fb = mel_filterbank.toarray() if hasattr(mel_filterbank, "toarray") else np.asarray(mel_filterbank)  # zaf.py:373
    fb = np.ascontiguousarray(fb, dtype=np.float64)
    return fb, fb.tobytes() if fb.ndim == 2 else b""


def _mel_plan(window_function, step_length, mel_filterbank, number_coefficients, route="fused", precision="auto"):
    w = _window64(window_function)
    fb, fb_key = _mel_ops.get(mel_filterbank, _prepare_filterbank)
    if fb.ndim != 2 or fb.shape[1] != len(w) // 2:
        raise ValueError(f"mel_filterbank must have shape (number_mels, window_length/2 = {len(w) // 2})")
    if route not in _MEL_ROUTES:
        raise ValueError(f"route must be one of {sorted(_MEL_ROUTES)}")
    if precision not in _MEL_PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_MEL_PRECISIONS)}")
    key = ("mel", len(w), int(step_length), int(number_coefficients), route, precision, w.tobytes(), fb_key)
    fresh = key not in _mel_plans._d
    plan = _mel_plans.get(key, w.ctypes.data, len(w), int(step_length), fb.ctypes.data, fb.shape[0],
                          int(number_coefficients))
    if fresh and route != "fused":
        _lib.check(_lib.lib().zafb_mel_plan_set_route(plan, _MEL_ROUTES[route]))
    if fresh:
        _lib.check(_lib.lib().zafb_mel_plan_set_precision(plan, _MEL_PRECISIONS[precision], w.ctypes.data))
    return plan, w, fb.shape[0]


def _mel_like(fn, audio_signal, window_function, step_length, mel_filterbank, ncoef, rows_of, layout, stream, route,
              precision):
    plan, w, n_mels = _mel_plan(window_function, step_length, mel_filterbank, ncoef, route, precision)
    lay = _layout_id(layout)
    rows = rows_of(n_mels)
    if isinstance(audio_signal, DeviceArray):
        x = audio_signal
        one, batch, ns = _device_signal(x)
        nt = stft_geometry(ns, len(w), step_length)[1]
        mem_shape = (batch, nt, rows) if lay == LAYOUT_FRAME_MAJOR else (batch, rows, nt)
        out = DeviceArray(mem_shape[1:] if one else mem_shape, np.float32, transposed=lay == LAYOUT_FRAME_MAJOR)
        _lib.check(getattr(_lib.lib(), fn)(plan, C.c_void_p(x.ptr), batch, ns, x.pitch, C.c_void_p(out.ptr), lay,
                                           _stream_ptr(stream)))
        return out
    x, one = _signal_batch(audio_signal)
    batch, ns = x.shape
    nt = stft_geometry(ns, len(w), step_length)[1]
    mem, view = _matrix_out(batch, rows, nt, np.float32, lay, one)
    _lib.check(getattr(_lib.lib(), fn.replace("_f32", "_host_f32"))(plan, x.ctypes.data, batch, ns, ns, mem.ctypes.data, lay))
    return view


def melspectrogram(audio_signal, window_function, step_length, mel_filterbank, *, layout="frame_major", stream=None,
                   route="fused", precision="auto"):
    """Mel spectrogram -- drop-in for ``zaf.melspectrogram`` (zaf.py:324-375): filterbank times the
    magnitude of STFT rows 1..N/2 (no DC, with Nyquist).  Returns (number_mels, number_times).
    ``route="tensor"`` applies the filterbank as the dense product of zaf.py:373 on the tcgen05 tensor
    cores (3xTF32, window_length 1024); the default fuses the banded filterbank into the STFT kernel.
    ``precision="float64"`` computes the spectrum and the filterbank sums in double precision on the GPU."""
    return _mel_like("zafb_melspectrogram_f32", audio_signal, window_function, step_length, mel_filterbank, 0,
                     lambda n_mels: n_mels, layout, stream, route, precision)


def mfcc(audio_signal, window_function, step_length, mel_filterbank, number_coefficients, *,
         layout="frame_major", stream=None, route="fused", precision="auto"):
    """MFCCs -- drop-in for ``zaf.mfcc`` (zaf.py:378-454): orthonormal DCT-II over the mel axis of
    ln(filterbank @ |STFT|^2 + eps), rows 1..number_coefficients.  Returns (number_coefficients, number_times).
    ``precision="auto"`` (default): fp32 kernels, and every frame whose quietest mel band lies more than 80 dB below its
    strongest bin -- the fp32 floor of the FFT, which the logarithm would amplify (purely tonal material) -- is recomputed
    in double precision by a second kernel on the same stream, so the 1e-5 bar holds for any signal.  ``"float64"``
    computes every frame in double precision, ``"float32"`` none (the dense tensor-core route is always fp32)."""
    ncoef = int(number_coefficients)
    return _mel_like("zafb_mfcc_f32", audio_signal, window_function, step_length, mel_filterbank, ncoef,
                     lambda n_mels: max(0, min(ncoef, n_mels - 1)), layout, stream, route, precision)


# ------------------------------------------------------------------ CQT
def _prepare_cqt_kernel(cqt_kernel):
    import scipy.sparse

    k = scipy.sparse.csr_matrix(cqt_kernel)
    k.sort_indices()
    data = np.ascontiguousarray(k.data, dtype=np.complex128)
    indptr = np.ascontiguousarray(k.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(k.indices, dtype=np.int32)
    return k.shape, data, indptr, indices, (data.tobytes(), indices.tobytes(), indptr.tobytes())


def _cqt_plan(cqt_kernel, step, route="fused"):
    (nf, fft_length), data, indptr, indices, op_key = _cqt_ops.get(cqt_kernel, _prepare_cqt_kernel)
    if route not in _MEL_ROUTES:
        raise ValueError(f"route must be one of {sorted(_MEL_ROUTES)}")
    key = ("cqt", nf, fft_length, int(step), route) + op_key
    fresh = key not in _cqt_plans._d
    plan = _cqt_plans.get(key, nf, fft_length, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, int(step))
    if fresh and route != "fused":
        _lib.check(_lib.lib().zafb_cqt_plan_set_route(plan, _MEL_ROUTES[route]))
    return plan, nf, fft_length


def _cqt_like(audio_signal, sampling_frequency, time_resolution, octave_resolution, cqt_kernel, layout, stream, route):
    step = round(sampling_frequency / time_resolution)  # zaf.py:603 (Python round-half-even)
    plan, nf, fft_length = _cqt_plan(cqt_kernel, step, route)
    lay = _layout_id(layout)
    rows = octave_resolution if octave_resolution else nf
    if isinstance(audio_signal, DeviceArray):
        x = audio_signal
        one, batch, ns = _device_signal(x)
        nt = cqt_geometry(ns, sampling_frequency, time_resolution, fft_length)[1]
        mem_shape = (batch, nt, rows) if lay == LAYOUT_FRAME_MAJOR else (batch, rows, nt)
        out = DeviceArray(mem_shape[1:] if one else mem_shape, np.float32, transposed=lay == LAYOUT_FRAME_MAJOR)
        _lib.check(_lib.lib().zafb_cqt_f32(plan, C.c_void_p(x.ptr), batch, ns, x.pitch, int(octave_resolution),
                                           C.c_void_p(out.ptr), lay, _stream_ptr(stream)))
        return out
    x, one = _signal_batch(audio_signal)
    batch, ns = x.shape
    nt = cqt_geometry(ns, sampling_frequency, time_resolution, fft_length)[1]
    mem, view = _matrix_out(batch, rows, nt, np.float32, lay, one)
    _lib.check(_lib.lib().zafb_cqt_host_f32(plan, x.ctypes.data, batch, ns, ns, int(octave_resolution), mem.ctypes.data,
                                            lay))
    return view


def cqtspectrogram(audio_signal, sampling_frequency, time_resolution, cqt_kernel, *, layout="frame_major",
                   stream=None, route="fused"):
    """Constant-Q spectrogram -- drop-in for ``zaf.cqtspectrogram`` (zaf.py:562-635).
    Returns (number_frequencies, number_times) float32.  ``route="tensor"`` applies the kernel as a dense 3xTF32
    product on the tcgen05 tensor cores (fft_length 32768) instead of the banded multiply-add fused into the FFT kernel."""
    return _cqt_like(audio_signal, sampling_frequency, time_resolution, 0, cqt_kernel, layout, stream, route)


def cqtchromagram(audio_signal, sampling_frequency, time_resolution, octave_resolution, cqt_kernel, *,
                  layout="frame_major", stream=None, route="fused"):
    """CQT chromagram -- drop-in for ``zaf.cqtchromagram`` (zaf.py:638-700): rows i::octave_resolution
    of the CQT spectrogram summed.  Returns (octave_resolution, number_times) float32."""
    return _cqt_like(audio_signal, sampling_frequency, time_resolution, int(octave_resolution), cqt_kernel, layout,
                     stream, route)

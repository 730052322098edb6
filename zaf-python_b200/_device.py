"""Minimal device-memory / stream / event objects on top of the C ABI (no PyTorch, no CuPy)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_state = {"device": None}


def device_count() -> int:
    n = C.c_int(0)
    try:
        rc = _lib.lib().zafb_device_count(C.byref(n))
    except _lib.ZafbError:
        return 0
    return n.value if rc == 0 else 0


def init(device: int = 0) -> None:
    """Select the CUDA device for this process (one process per GPU)."""
    _lib.check(_lib.lib().zafb_init(int(device)))
    _state["device"] = int(device)


def ensure_init() -> None:
    if _state["device"] is None:
        init(0)


def synchronize() -> None:
    _lib.check(_lib.lib().zafb_device_sync())


def host_copy_bytes():
    """(h2d, d2h) bytes the host-buffer pipelines have moved over the link so far."""
    import ctypes as C
    a, b = C.c_int64(0), C.c_int64(0)
    _lib.check(_lib.lib().zafb_host_copy_bytes(C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def launch_count() -> int:
    return int(_lib.lib().zafb_launch_count())


class DeviceArray:
    """A typed, shaped view of a device allocation.  ``strides_view`` records whether the logical
    array is the transpose of the last two axes of the memory (frame-major transforms)."""

    def __init__(self, shape, dtype, ptr=None, owner=None, transposed=False, cols=None):
        ensure_init()
        self.mem_shape = tuple(int(s) for s in shape)  # shape of the memory, C order
        # rows padded to an aligned pitch: the logical last axis has `cols` <= mem_shape[-1] elements
        self.cols = None if cols is None or int(cols) == self.mem_shape[-1] else int(cols)
        self.dtype = np.dtype(dtype)
        self.transposed = bool(transposed)
        self.nbytes = int(np.prod(self.mem_shape, dtype=np.int64)) * self.dtype.itemsize
        self._owner = owner
        if ptr is None:
            p = C.c_void_p()
            _lib.check(_lib.lib().zafb_malloc(C.byref(p), self.nbytes))
            self.ptr = p.value or 0
            self._owns = True
        else:
            self.ptr = int(ptr)
            self._owns = False

    @property
    def shape(self):
        if self.transposed:
            return self.mem_shape[:-2] + (self.mem_shape[-1], self.mem_shape[-2])
        if self.cols is not None:
            return self.mem_shape[:-1] + (self.cols,)
        return self.mem_shape

    @property
    def pitch(self):
        """Elements between consecutive rows of the memory (>= the logical row length)."""
        return self.mem_shape[-1]

    def free(self):
        if self._owns and self.ptr:
            _lib.lib().zafb_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def to_host(self, out=None, stream=None):
        """Copy to a NumPy array of the logical shape (a transposed *view* of the copied memory
        for frame-major results -- no data movement on the host).  ``out=``: a C-contiguous array of
        this dtype holding either the logical rows (padding columns are skipped) or the whole padded memory."""
        sp = stream.ptr if stream else None
        item = self.dtype.itemsize
        if out is not None:
            if not isinstance(out, np.ndarray) or out.dtype != self.dtype or not out.flags.c_contiguous:
                raise ValueError(f"out must be a C-contiguous NumPy array of dtype {self.dtype}")
            logical = int(np.prod(self.mem_shape[:-1], dtype=np.int64)) * (self.cols or self.mem_shape[-1]) * item
            if out.nbytes not in (logical, self.nbytes):
                raise ValueError(f"out holds {out.nbytes} bytes; the array has {logical} (memory: {self.nbytes})")
        packed = self.cols is not None and (out is None or out.nbytes != self.nbytes)
        if packed:  # padded rows: copy only the logical columns
            host = np.empty(self.mem_shape[:-1] + (self.cols,), dtype=self.dtype) if out is None else out
            rows = int(np.prod(self.mem_shape[:-1], dtype=np.int64))
            _lib.check(_lib.lib().zafb_memcpy2d(host.ctypes.data, self.cols * item, C.c_void_p(self.ptr),
                                                self.mem_shape[-1] * item, self.cols * item, rows, 1, sp))
            if stream is None:
                synchronize()
            return host.reshape(self.mem_shape[:-1] + (self.cols,))
        host = np.empty(self.mem_shape, dtype=self.dtype) if out is None else out.reshape(self.mem_shape)
        _lib.check(_lib.lib().zafb_memcpy_d2h(host.ctypes.data, C.c_void_p(self.ptr), self.nbytes, sp))
        if stream is None:
            synchronize()
        if self.transposed:
            return np.swapaxes(host, -1, -2)
        return host[..., :self.cols] if self.cols is not None else host

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, transposed={self.transposed}, pitch={self.pitch})"


def to_device(array, dtype=None, stream=None) -> DeviceArray:
    """Host array -> DeviceArray.  A batch of float32 signals with an odd number of samples per row gets an even row
    pitch on the device (one padding column, never read as signal) so that every row stays 8-byte aligned for the
    vectorised kernels."""
    a = np.ascontiguousarray(array, dtype=dtype)
    if a.ndim == 2 and a.dtype == np.float32 and a.shape[0] > 1 and a.shape[1] % 2 == 1:
        pitch = a.shape[1] + 1
        d = DeviceArray((a.shape[0], pitch), a.dtype, cols=a.shape[1])
        _lib.check(_lib.lib().zafb_memcpy2d(C.c_void_p(d.ptr), pitch * 4, a.ctypes.data, a.shape[1] * 4, a.shape[1] * 4,
                                            a.shape[0], 0, stream.ptr if stream else None))
        if stream is None:
            synchronize()
        return d
    d = DeviceArray(a.shape, a.dtype)
    if a.nbytes:
        _lib.check(_lib.lib().zafb_memcpy_h2d(C.c_void_p(d.ptr), a.ctypes.data, a.nbytes,
                                              stream.ptr if stream else None))
        if stream is None:
            synchronize()
    return d


def empty(shape, dtype) -> DeviceArray:
    return DeviceArray(shape, dtype)


class PinnedArray:
    """Page-locked host memory exposed as a NumPy array (``.array``)."""

    def __init__(self, shape, dtype):
        ensure_init()
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64))
        self.nbytes = n * self.dtype.itemsize
        p = C.c_void_p()
        _lib.check(_lib.lib().zafb_host_alloc(C.byref(p), max(self.nbytes, 1)))
        self.ptr = p.value
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=n).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            _lib.lib().zafb_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Stream:
    def __init__(self):
        ensure_init()
        p = C.c_void_p()
        _lib.check(_lib.lib().zafb_stream_create(C.byref(p)))
        self.ptr = p

    def synchronize(self):
        _lib.check(_lib.lib().zafb_stream_sync(self.ptr))

    def wait_event(self, event):
        """Work queued on this stream from now on starts only after ``event`` (recorded on another stream) completed."""
        _lib.check(_lib.lib().zafb_stream_wait_event(self.ptr, event.ptr))

    def __del__(self):
        try:
            _lib.lib().zafb_stream_destroy(self.ptr)
        except Exception:
            pass


class Event:
    def __init__(self):
        ensure_init()
        p = C.c_void_p()
        _lib.check(_lib.lib().zafb_event_create(C.byref(p)))
        self.ptr = p

    def record(self, stream=None):
        _lib.check(_lib.lib().zafb_event_record(self.ptr, stream.ptr if stream else None))

    def synchronize(self):
        _lib.check(_lib.lib().zafb_event_sync(self.ptr))

    def elapsed_ms(self, later: "Event") -> float:
        ms = C.c_float(0)
        _lib.check(_lib.lib().zafb_event_elapsed_ms(self.ptr, later.ptr, C.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            _lib.lib().zafb_event_destroy(self.ptr)
        except Exception:
            pass

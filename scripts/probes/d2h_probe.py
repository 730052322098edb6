#!/usr/bin/env python
"""Where does the half-spectrum host path lose time?  Pinned D2H variants of the cfg-2 result (1024 clips x 939 frames):
linear, strided rows of 8200 / 8192 bytes at a 16384-byte pitch, chunked over three streams like the host pipeline, with
and without the concurrent H2D of the input and the host fill threads."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import zaf_python_b200 as zaf  # noqa: E402

zaf.init(0)
lib = zaf._lib.lib()
clips, nt, n = int(os.environ.get("CLIPS", "512")), 939, 2048
rows = clips * nt
dev = zaf.empty((rows, n), np.complex64)
pin = zaf.PinnedArray((rows, n), np.complex64)
pin.array[:] = 0
streams = [zaf.Stream() for _ in range(3)]


def timed(label, fn, nbytes, reps=3):
    best = 1e9
    for _ in range(reps):
        zaf.synchronize()
        t0 = time.perf_counter()
        fn()
        zaf.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{label:62s} {best * 1e3:8.1f} ms  {nbytes / best / 1e9:6.1f} GB/s", flush=True)


half = rows * (n // 2 + 1) * 8
timed("linear D2H, same bytes as the half spectrum", lambda: zaf._lib.check(lib.zafb_memcpy_d2h(pin.array.ctypes.data, C.c_void_p(dev.ptr), half, None)), half)
for width in (8200, 8192, 8224):
    timed(f"2D D2H, {width}-byte rows at pitch 16384 (one call)",
          lambda: zaf._lib.check(lib.zafb_memcpy2d(pin.array.ctypes.data, 16384, C.c_void_p(dev.ptr), 16384, width, rows, 1, None)), rows * width)
timed("2D D2H, compact device rows (pitch 8200) -> host pitch 16384",
      lambda: zaf._lib.check(lib.zafb_memcpy2d(pin.array.ctypes.data, 16384, C.c_void_p(dev.ptr), 8200, 8200, rows, 1, None)), half)
for chunk_clips in (1, 4, 16):
    def chunked():
        for i, c0 in enumerate(range(0, clips, chunk_clips)):
            r0, nr = c0 * nt, min(chunk_clips, clips - c0) * nt
            zaf._lib.check(lib.zafb_memcpy2d(pin.array.ctypes.data + r0 * 16384, 16384, C.c_void_p(dev.ptr + r0 * 16384), 16384, 8200, nr, 1,
                                             streams[i % 3].ptr))
    timed(f"2D D2H 8200/16384 in chunks of {chunk_clips} clip(s) on 3 streams", chunked, half)
# the host fill alone (no copies): 16 threads through the library's own entry point is not exposed; time one thread
t0 = time.perf_counter()
zaf._lib.check(lib.zafb_host_mirror_fill(C.c_void_p(pin.array.ctypes.data), rows // 16, n))
dt = time.perf_counter() - t0
print(f"host mirror fill, ONE thread, 1/16 of the frames: {dt * 1e3:.1f} ms -> {rows // 16 * (n // 2 - 1) * 16 / dt / 1e9:.1f} GB/s read+write per thread")
import threading
def fill_part(i, parts):
    lo, hi = rows * i // parts, rows * (i + 1) // parts
    lib.zafb_host_mirror_fill(C.c_void_p(pin.array.ctypes.data + lo * 16384), hi - lo, n)
for parts in (4, 8, 16):
    th = [threading.Thread(target=fill_part, args=(i, parts)) for i in range(parts)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"host mirror fill, {parts} threads, all frames: {dt * 1e3:.1f} ms -> {rows * (n // 2 - 1) * 16 / dt / 1e9:.1f} GB/s read+write")

#!/usr/bin/env python
"""HBM probe through the library's C ABI: write-only (memset) and read+write (device-to-device copy)
bandwidth, to put the write-dominated STFT kernel (16 KB written per 2 KB read) next to the right ceiling."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zaf_python_b200 as zaf  # noqa: E402

zaf.init(0)
lib, C = zaf._lib.lib(), zaf._lib.C
n = 8 << 30
a = zaf.empty((n,), np.uint8)
b = zaf.empty((n,), np.uint8)
st = zaf.Stream()
e0, e1 = zaf.Event(), zaf.Event()


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    st.synchronize()
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    e1.synchronize()
    return e0.elapsed_ms(e1) / reps


ms = timed(lambda: zaf._lib.check(lib.zafb_memset(C.c_void_p(a.ptr), 1, n, st.ptr)))
print(json.dumps({"probe": "memset_write_only", "gb": n / 1e9, "ms": ms, "gbs": n / ms / 1e6}))
ms = timed(lambda: zaf._lib.check(lib.zafb_memcpy_d2d(C.c_void_p(b.ptr), C.c_void_p(a.ptr), n, st.ptr)))
print(json.dumps({"probe": "d2d_copy_read+write", "gb_moved": 2 * n / 1e9, "ms": ms, "gbs": 2 * n / ms / 1e6}))

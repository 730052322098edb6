#!/usr/bin/env python
"""Per-stage device timing of the one-sided stft -> |X| -> min-ratio mask -> multiply -> istft chain (cfg-2 batch).
usage: python scripts/probes/chain_probe.py [clips]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import zaf_python_b200 as zaf  # noqa: E402
from bench_configs import device_batch, hamming_periodic  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
zaf.init(0)
lib = zaf._lib.lib()
NS, N, HOP = 480000, 2048, 512
w = hamming_periodic(N)
plan, _ = zaf._stft_plan(w, HOP)
xd, _ = device_batch(clips, NS, 1)
nt = zaf.stft_geometry(NS, N, HOP)[1]
k = N // 2 + 1
pitch = (k + 3) & ~3
half = zaf.empty((clips, nt, pitch), np.complex64)
mag = zaf.empty((clips, nt, pitch), np.float32)
swp = zaf.empty((clips, nt, pitch), np.float32)
ylen = zaf.istft_geometry(N, nt, HOP)[2]
yd = zaf.empty((clips, ylen), np.float32)
pair = nt * pitch * 4
st = zaf.Stream()
chk = zaf._lib.check
stages = [
    ("stft_onesided", lambda: chk(lib.zafb_stft_onesided_f32(plan, C.c_void_p(xd.ptr), clips, NS, NS, C.c_void_p(half.ptr), pitch, st.ptr))),
    ("spec_abs", lambda: chk(lib.zafb_spec_abs_f32(C.c_void_p(half.ptr), clips, pitch, nt, 0, pitch, C.c_void_p(mag.ptr), st.ptr))),
    ("swap (2 x memcpy2d)", lambda: (chk(lib.zafb_memcpy2d(C.c_void_p(swp.ptr), 2 * pair, C.c_void_p(mag.ptr + pair), 2 * pair, pair, clips // 2, 2, st.ptr)),
                                      chk(lib.zafb_memcpy2d(C.c_void_p(swp.ptr + pair), 2 * pair, C.c_void_p(mag.ptr), 2 * pair, pair, clips // 2, 2, st.ptr)))),
    ("ratio_min", lambda: chk(lib.zafb_ratio_min_f32(C.c_void_p(mag.ptr), C.c_void_p(swp.ptr), clips * nt * pitch, C.c_void_p(mag.ptr), st.ptr))),
    ("spec_mask", lambda: chk(lib.zafb_spec_mask_f32(C.c_void_p(half.ptr), clips, pitch, nt, 0, C.c_void_p(mag.ptr), pitch, C.c_void_p(half.ptr), st.ptr))),
    ("istft_onesided", lambda: chk(lib.zafb_istft_onesided_f32(plan, C.c_void_p(half.ptr), clips, nt, pitch, C.c_void_p(yd.ptr), ylen, st.ptr))),
]
chk(lib.zafb_memset(C.c_void_p(half.ptr), 0, half.nbytes, st.ptr))
for _ in range(2):
    for _, f in stages:
        f()
st.synchronize()
ev = [zaf.Event() for _ in range(len(stages) + 1)]
tot = [0.0] * len(stages)
for _ in range(3):
    ev[0].record(st)
    for i, (_, f) in enumerate(stages):
        f()
        ev[i + 1].record(st)
    ev[-1].synchronize()
    for i in range(len(stages)):
        tot[i] += ev[i].elapsed_ms(ev[i + 1]) / 3
for (name, _), t in zip(stages, tot):
    print(f"{name:22s} {t:7.3f} ms")
print(f"{'total':22s} {sum(tot):7.3f} ms")

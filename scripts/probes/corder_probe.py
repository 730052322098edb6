#!/usr/bin/env python
"""Device-resident timing of the transforms in the reference's C-order memory (layout = BIN_MAJOR) on BASELINE shapes
(a fraction of the batch): stft, istft (cfg 2), mdct, imdct (cfg 4).  Environment switches select the route
(ZAFB_STFT_BM_DIRECT, ZAFB_MDCT_BM_DIRECT, ZAFB_IMDCT_BM_DIRECT, ZAFB_TRANSPOSE_CHUNK_MB, ...).
usage: python scripts/probes/corder_probe.py [scale] [mdct]   ("mdct": the cfg-4 half only)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import zaf_python_b200 as zaf  # noqa: E402
from bench_configs import device_batch, hamming_periodic, kbd, peak, timeit  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
zaf.init(0)
lib = zaf._lib.lib()
tag = {k: v for k, v in os.environ.items() if k.startswith("ZAFB_")}


def emit(name, ms, nbytes, launches):
    print(json.dumps({"transform": name, "ms": round(ms, 3), "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / peak(), 3),
                      "launches": launches, "env": tag}), flush=True)


mdct_only = len(sys.argv) > 2 and sys.argv[2] == "mdct"
clips, ns, n, hop = (1 if mdct_only else int(1024 * scale)), 480000, 2048, 512
w = hamming_periodic(n)
xd, _ = device_batch(clips, ns, 1)
nt = zaf.stft_geometry(ns, n, hop)[1]
spec = zaf.empty((clips, n, nt), np.complex64)
plan, _ = zaf._stft_plan(w, hop)
ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 1, s.ptr)), 5)
emit("stft bin_major", ms, clips * ns * 4 + clips * nt * n * 8, nl)
ylen = zaf.istft_geometry(n, nt, hop)[2]
yd = zaf.empty((clips, ylen), np.float32)
ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_istft_f32(plan, C.c_void_p(spec.ptr), clips, nt, 1, C.c_void_p(yd.ptr), ylen, s.ptr)), 5)
emit("istft bin_major", ms, clips * nt * n * 8 + clips * ylen * 4, nl)
xd.free(), spec.free(), yd.free()

clips, ns, n = int(2048 * scale), 1323000, 2048
w = kbd(n)
xd, _ = device_batch(clips, ns, 2)
m, nt, _ = zaf.mdct_geometry(ns, n)
plan, _ = zaf._mdct_plan(w)
spec = zaf.empty((clips, m, nt), np.float32)
ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mdct_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 1, s.ptr)), 5)
emit("mdct bin_major", ms, clips * ns * 4 + clips * nt * m * 4, nl)
ylen = zaf.imdct_geometry(m, nt)[1]
pitch = (ylen + 1) & ~1
yd = zaf.empty((clips, pitch), np.float32)
ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_imdct_f32(plan, C.c_void_p(spec.ptr), clips, nt, 1, C.c_void_p(yd.ptr), pitch, s.ptr)), 5)
emit("imdct bin_major", ms, clips * nt * m * 4 + clips * ylen * 4, nl)

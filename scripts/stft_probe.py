#!/usr/bin/env python
"""Probe: device-resident STFT / ISTFT time for any window length -- python scripts/stft_probe.py N:HOP:NS:CLIPS [...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import zaf_python_b200 as zaf  # noqa: E402
from bench_configs import device_batch, hamming_periodic, peak, timeit  # noqa: E402

zaf.init(0)
lib, C = zaf._lib.lib(), zaf._lib.C
for spec_ in sys.argv[1:]:
    n, hop, ns, clips = (int(v) for v in spec_.split(":"))
    w = hamming_periodic(n)
    plan, _ = zaf._stft_plan(w, hop)
    xd, _ = device_batch(clips, ns, 1)
    nt = zaf.stft_geometry(ns, n, hop)[1]
    spec = zaf.empty((clips, nt, n), np.complex64)
    ylen = zaf.istft_geometry(n, nt, hop)[2]
    yd = zaf.empty((clips, ylen + 1), np.float32)
    for layout in (0, 1):
        def f(stream):
            zaf._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), layout, stream.ptr))

        def g(stream):
            zaf._lib.check(lib.zafb_istft_f32(plan, C.c_void_p(spec.ptr), clips, nt, layout, C.c_void_p(yd.ptr), (ylen + 1) & ~1, stream.ptr))
        ms, _, nl = timeit(f, 10)
        by = clips * ns * 4 + clips * nt * n * 8
        ms2, _, nl2 = timeit(g, 10)
        print(f"N={n} hop={hop} ns={ns} clips={clips} nt={nt} layout={layout}: stft {ms:.3f} ms ({by / ms / 1e6 / peak():.3f} of HBM peak, {nl} launches)"
              f"  istft {ms2:.3f} ms ({by / ms2 / 1e6 / peak():.3f}, {nl2} launches)", flush=True)
    xd.free(); spec.free(); yd.free()

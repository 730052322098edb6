#!/usr/bin/env python
"""Probe: device-resident MDCT / IMDCT time for any window length -- python scripts/mdct_probe.py N:NS:CLIPS [...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import zaf_python_b200 as zaf  # noqa: E402
from bench_configs import device_batch, kbd, peak, timeit  # noqa: E402

zaf.init(0)
lib, C = zaf._lib.lib(), zaf._lib.C
for spec_ in sys.argv[1:]:
    n, ns, clips = (int(v) for v in spec_.split(":"))
    w = kbd(n)
    xd, _ = device_batch(clips, ns, 1)
    m, nt, _ = zaf.mdct_geometry(ns, n)
    plan, _ = zaf._mdct_plan(w)
    spec = zaf.empty((clips, nt, m), np.float32)
    ylen = zaf.imdct_geometry(m, nt)[1]
    pitch = (ylen + 1) & ~1
    yd = zaf.empty((clips, pitch), np.float32)
    for layout in (0, 1):
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mdct_f32(
            plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), layout, s.ptr)), 10)
        ms2, _, nl2 = timeit(lambda s: zaf._lib.check(lib.zafb_imdct_f32(
            plan, C.c_void_p(spec.ptr), clips, nt, layout, C.c_void_p(yd.ptr), pitch, s.ptr)), 10)
        by = clips * ns * 4 + clips * nt * m * 4
        print(f"N={n} ns={ns} clips={clips} nt={nt} layout={layout}: mdct {ms:.3f} ms ({by / ms / 1e6 / peak():.3f} of HBM peak, {nl} launches)"
              f"  imdct {ms2:.3f} ms ({by / ms2 / 1e6 / peak():.3f}, {nl2} launches)", flush=True)
    xd.free(); spec.free(); yd.free()

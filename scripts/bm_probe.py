#!/usr/bin/env python
"""Probe: BIN_MAJOR STFT time against the frame count per clip (row alignment) -- python scripts/bm_probe.py NS [NS ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import zaf_python_b200 as zaf  # noqa: E402
from bench_configs import device_batch, hamming_periodic, timeit  # noqa: E402

zaf.init(0)
clips, n, hop = 1024, 2048, 512
w = hamming_periodic(n)
plan, _ = zaf._stft_plan(w, hop)
lib, C = zaf._lib.lib(), zaf._lib.C
for ns in [int(a) for a in sys.argv[1:]]:
    xd, _ = device_batch(clips, ns, 1)
    nt = zaf.stft_geometry(ns, n, hop)[1]
    spec = zaf.empty((clips, nt, n), np.complex64)
    for layout in (0, 1):
        def f(stream):
            zaf._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), layout, stream.ptr))
        ms, _, nl = timeit(f, 10)
        print(f"ns={ns} nt={nt} (nt mod 4 = {nt % 4}) layout={layout} env={ {k: v for k, v in os.environ.items() if k.startswith('ZAFB_')} }: {ms:.3f} ms", flush=True)
    xd.free()
    spec.free()

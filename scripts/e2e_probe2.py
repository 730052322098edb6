#!/usr/bin/env python
"""Probe: end-to-end zaf.stft on the cfg-2 batch (host in, pinned host out) under the current ZAFB_* environment."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zaf_python_b200 as zaf  # noqa: E402

zaf.init(0)
clips, ns, n, hop = 1024, 480000, 2048, 512
w = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)
nt = zaf.stft_geometry(ns, n, hop)[1]
pin_x = zaf.PinnedArray((clips, ns), np.float32)
pin_x.array[:] = np.random.default_rng(1).uniform(-1, 1, (32, ns)).astype(np.float32)[np.arange(clips) % 32]
pin_out = zaf.PinnedArray((clips, nt, n), np.complex64)
for layout in sys.argv[1:] or ["frame_major"]:
    out = pin_out.array if layout == "frame_major" else pin_out.array.reshape(clips, n, nt)
    zaf.stft(pin_x.array, w, hop, out=out, layout=layout)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        zaf.stft(pin_x.array, w, hop, out=out, layout=layout)
        best = min(best, time.perf_counter() - t0)
    env = {k: v for k, v in os.environ.items() if k.startswith("ZAFB_")}
    print(f"{layout} {env}: best {best * 1e3:.1f} ms  {clips * nt / best:.4g} frames/s", flush=True)

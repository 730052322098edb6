#!/usr/bin/env python
"""End-to-end (host buffers) timing of the drop-in calls with the pipeline trace on, for a few
pipeline chunk sizes.  usage: ZAFB_TRACE=1 python scripts/e2e_probe.py [--clips 1024]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zaf_python_b200 as zaf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--chunks", default="16,64,256")
    args = ap.parse_args()
    zaf.init(0)
    n, hop, ns = 2048, 512, 480000
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)
    nt = zaf.stft_geometry(ns, n, hop)[1]
    pin_x = zaf.PinnedArray((args.clips, ns), np.float32)
    pin_x.array[:] = np.random.default_rng(0).uniform(-1, 1, (1, ns)).astype(np.float32)
    pin_out = zaf.PinnedArray((args.clips, nt, n), np.complex64)
    for mb in [int(c) for c in args.chunks.split(",")]:
        os.environ["ZAFB_PIPE_CHUNK_MB"] = str(mb)
        for rep in range(3):
            t0 = time.perf_counter()
            zaf.stft(pin_x.array, w, hop, out=pin_out.array)
            dt = time.perf_counter() - t0
            print(f"chunk {mb} MB rep {rep}: {dt * 1e3:.1f} ms  {args.clips * nt / dt:.4g} frames/s  "
                  f"D2H {pin_out.nbytes / dt / 1e9:.1f} GB/s", flush=True)
    # the default call (no out=): the result comes from the pinned pool -- first call locks the pages, later calls reuse them
    for rep in range(3):
        t0 = time.perf_counter()
        spec = zaf.stft(pin_x.array[:256], w, hop)
        dt = time.perf_counter() - t0
        print(f"default call (pooled pinned result), 256 clips, rep {rep}: {dt * 1e3:.1f} ms  {256 * nt / dt:.4g} frames/s",
              flush=True)
        del spec
    os.environ["ZAFB_PINNED_POOL_MB"] = "0"
    zaf._pinned._pool.cap = 0
    t0 = time.perf_counter()
    spec = zaf.stft(pin_x.array[:128], w, hop)
    dt = time.perf_counter() - t0
    print(f"pageable result (pool off), 128 clips: {dt * 1e3:.1f} ms  {128 * nt / dt:.4g} frames/s", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Host<->device link probe: what the end-to-end (host buffers) numbers of bench.py are bounded by.

Times pinned-memory copies through the library's own C ABI (zafb_memcpy_*): D2H and H2D alone at
several sizes, D2H split over two streams, and D2H with a concurrent H2D.  One JSON line per row.

    python scripts/pcie_probe.py [--out FILE]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zaf_python_b200 as zaf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--mb", type=int, default=2048)
    args = ap.parse_args()
    zaf.init(0)
    lib, C = zaf._lib.lib(), zaf._lib.C
    nbytes = args.mb << 20
    pin = zaf.PinnedArray((nbytes,), np.uint8)
    pin2 = zaf.PinnedArray((nbytes,), np.uint8)
    pin.array[:] = 1
    pin2.array[:] = 2
    dev = zaf.empty((nbytes,), np.uint8)
    dev2 = zaf.empty((nbytes,), np.uint8)
    s0, s1 = zaf.Stream(), zaf.Stream()
    lines = []

    def emit(d):
        lines.append(d)
        print(json.dumps(d), flush=True)

    def d2h(dst, src, n, st, off=0):
        zaf._lib.check(lib.zafb_memcpy_d2h(C.c_void_p(dst.ptr + off), C.c_void_p(src.ptr + off), n, st.ptr))

    def h2d(dst, src, n, st, off=0):
        zaf._lib.check(lib.zafb_memcpy_h2d(C.c_void_p(dst.ptr + off), C.c_void_p(src.ptr + off), n, st.ptr))

    def wall(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            zaf.synchronize()
            t0 = time.perf_counter()
            fn()
            zaf.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    for chunk_mb in (16, 64, 256, args.mb):
        chunk = chunk_mb << 20

        def run_d2h():
            for off in range(0, nbytes, chunk):
                d2h(pin, dev, chunk, s0, off)

        def run_h2d():
            for off in range(0, nbytes, chunk):
                h2d(dev, pin, chunk, s0, off)

        t = wall(run_d2h)
        emit({"probe": "d2h", "chunk_mb": chunk_mb, "total_mb": args.mb, "gbs": nbytes / t / 1e9})
        t = wall(run_h2d)
        emit({"probe": "h2d", "chunk_mb": chunk_mb, "total_mb": args.mb, "gbs": nbytes / t / 1e9})

    chunk = 256 << 20

    def run_d2h_2streams():
        for i, off in enumerate(range(0, nbytes, chunk)):
            d2h(pin, dev, chunk, s0 if i % 2 == 0 else s1, off)

    t = wall(run_d2h_2streams)
    emit({"probe": "d2h_two_streams", "chunk_mb": 256, "total_mb": args.mb, "gbs": nbytes / t / 1e9})

    def run_duplex():
        for off in range(0, nbytes, chunk):
            d2h(pin, dev, chunk, s0, off)
            h2d(dev2, pin2, chunk, s1, off)

    t = wall(run_duplex)
    emit({"probe": "d2h+h2d_concurrent", "chunk_mb": 256, "total_mb": args.mb,
          "gbs_each_direction": nbytes / t / 1e9})

    # host memory bandwidth seen by one thread (what a host-side copy of the result would cost)
    a = np.empty(nbytes, np.uint8)
    t0 = time.perf_counter()
    a[:] = pin.array
    emit({"probe": "host_memcpy_1thread", "gbs": nbytes / (time.perf_counter() - t0) / 1e9})

    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["nvidia-smi", "-q", "-d", "PCIE"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
        except Exception as exc:  # noqa: BLE001
            txt = f"{cmd}: {exc}"
        lines.append({"cmd": " ".join(cmd), "out": txt[:6000]})
    if args.out:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()

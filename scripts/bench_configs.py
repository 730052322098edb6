#!/usr/bin/env python
"""Device-resident timing of every BASELINE.json config, one JSON line each.

bench.py carries the headline metric (cfg 2 forward STFT).  This script times the other transforms
of the hot path on their BASELINE shapes with inputs already in HBM, CUDA events on the launching
stream, and reports algorithmic bytes / launch duration against the measured HBM peak.

    python scripts/bench_configs.py [--only stft,istft,mdct,imdct,mel,mfcc,cqt,dct] [--scale 1.0] [--out FILE]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zaf_python_b200 as zaf  # noqa: E402


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def hamming_periodic(n):
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)


def kbd(n, alpha=5.0):
    k = np.kaiser(n // 2 + 1, np.pi * alpha)
    half = np.sqrt(np.cumsum(k[: n // 2]) / np.sum(k))
    return np.concatenate([half, half[::-1]])


def device_batch(clips, ns, seed, distinct=32):
    """(clips, ns) float32 on the device, built from `distinct` seeded clips tiled over the batch."""
    rng = np.random.default_rng(seed)
    distinct = min(distinct, clips)
    host = rng.uniform(-1, 1, (distinct, ns)).astype(np.float32)
    d = zaf.empty((clips, ns), np.float32)
    lib, C = zaf._lib.lib(), zaf._lib.C
    for c0 in range(0, clips, distinct):
        n = min(distinct, clips - c0)
        zaf._lib.check(lib.zafb_memcpy_h2d(C.c_void_p(d.ptr + c0 * ns * 4), host.ctypes.data, n * ns * 4, None))
    zaf.synchronize()
    return d, host


def timeit(fn, steps, warmup=3):
    stream = zaf.Stream()
    out = None
    for _ in range(warmup):
        out = fn(stream)
    stream.synchronize()
    e0, e1 = zaf.Event(), zaf.Event()
    l0 = zaf.launch_count()
    e0.record(stream)
    for _ in range(steps):
        o = fn(stream)
        if o is not None and o is not out and hasattr(o, "free"):
            o.free()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_ms(e1) / steps, out, (zaf.launch_count() - l0) // steps


def line(name, cfg, frames, unit, ms, algo_bytes, launches, note=""):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    return {"transform": name, "config": cfg, "units": frames, "unit": unit, "ms_per_step": ms,
            "units_per_sec": frames / (ms * 1e-3), "algorithmic_bytes": int(algo_bytes), "achieved_gbs": gbs,
            "hbm_peak_gbs": peak(), "hbm_frac": gbs / peak(), "launches_per_step": int(launches), "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg1,stft,stftbin,istft,stft1024,stft512,mdct,imdct,mdct1024,mel,mfcc,meltc,melf64,mel2048,cqt,cqttc,dct")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of each BASELINE batch")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    only = set(args.only.split(","))
    zaf.init(0)
    out_lines = []

    def emit(d):
        out_lines.append(d)
        print(json.dumps(d), flush=True)

    # ---- cfg 1: the reference's own CPU-runnable case, ONE clip through the drop-in host call (wall clock, H2D + D2H inside)
    if "cfg1" in only:
        import time

        ns, n, hop = 480000, 2048, 1024
        w = hamming_periodic(n)
        x = np.random.default_rng(20261017 + 1).uniform(-1, 1, ns).astype(np.float32)
        nt = zaf.stft_geometry(ns, n, hop)[1]
        best = 1e9
        for _ in range(30):
            t0 = time.perf_counter()
            spec = zaf.stft(x, w, hop)
            best = min(best, time.perf_counter() - t0)
        assert spec.shape == (n, nt)
        emit(line("stft-cfg1-host-call", "cfg1: 1 clip x 10 s @ 48 kHz, N=2048 hop=1024, zaf.stft(x, w, hop) with NumPy in/out", nt,
                  "frames", best * 1e3, ns * 4 + nt * n * 8, 1, "end-to-end latency of one drop-in call (best of 30), not a device-resident rate"))

    # ---- cfg 2: stft + istft, 1024 clips x 10 s @ 48 kHz, N = 2048, hop = 512
    if only & {"stft", "istft", "stftbin"}:
        clips, ns, n, hop = max(1, int(1024 * args.scale)), 480000, 2048, 512
        w = hamming_periodic(n)
        xd, _ = device_batch(clips, ns, 20261017 + 2)
        nt = zaf.stft_geometry(ns, n, hop)[1]
        spec = zaf.empty((clips, nt, n), np.complex64)
        spec.transposed = True
        plan, _ = zaf._stft_plan(w, hop)
        lib, C = zaf._lib.lib(), zaf._lib.C
        cfg = f"cfg2: {clips} clips x 10 s @ 48 kHz, N=2048 hop=512"

        def f_stft(stream):
            zaf._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 0, stream.ptr))

        ms, _, nl = timeit(f_stft, args.steps)
        if "stft" in only:
            emit(line("stft", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * n * 8, nl))
        if "stftbin" in only:  # the reference's C-order memory (bin-major)
            def f_stft_bin(stream):
                zaf._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 1, stream.ptr))

            ms, _, nl = timeit(f_stft_bin, max(2, args.steps // 5))
            emit(line("stft-bin-major", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * n * 8, nl,
                      "layout = BIN_MAJOR: [clip][bin][frame], the reference's C order"))
            f_stft(stream=zaf.Stream())  # leave a frame-major spectrum behind for the istft leg
            zaf.synchronize()
        if "istft" in only:
            ylen = zaf.istft_geometry(n, nt, hop)[2]
            yd = zaf.empty((clips, ylen), np.float32)

            def f_istft(stream):
                zaf._lib.check(lib.zafb_istft_f32(plan, C.c_void_p(spec.ptr), clips, nt, 0, C.c_void_p(yd.ptr), ylen, stream.ptr))

            ms, _, nl = timeit(f_istft, args.steps)
            emit(line("istft", cfg, clips * nt, "frames", ms, clips * nt * n * 8 + clips * ylen * 4, nl))
            yd.free()
        xd.free()
        spec.free()

    # ---- cfg 3 shape through plain stft / istft (N = 1024, hop = 256): the second window length with warp kernels
    if "stft1024" in only:
        clips, ns, n, hop = max(1, int(4096 * args.scale)), 80000, 1024, 256
        w = hamming_periodic(n)
        xd, _ = device_batch(clips, ns, 20261017 + 3)
        nt = zaf.stft_geometry(ns, n, hop)[1]
        spec = zaf.empty((clips, nt, n), np.complex64)
        plan, _ = zaf._stft_plan(w, hop)
        lib, C = zaf._lib.lib(), zaf._lib.C
        cfg = f"cfg3 shape: {clips} clips x 5 s @ 16 kHz, N=1024 hop=256"
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_stft_f32(
            plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 0, s.ptr)), args.steps)
        emit(line("stft-1024", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * n * 8, nl))
        ylen = zaf.istft_geometry(n, nt, hop)[2]
        yd = zaf.empty((clips, ylen), np.float32)
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_istft_f32(
            plan, C.c_void_p(spec.ptr), clips, nt, 0, C.c_void_p(yd.ptr), ylen, s.ptr)), args.steps)
        emit(line("istft-1024", cfg, clips * nt, "frames", ms, clips * nt * n * 8 + clips * ylen * 4, nl))
        xd.free(), spec.free(), yd.free()

    # ---- speech shape (N = 512, hop = 128 at 16 kHz): the third window length with warp kernels
    if "stft512" in only:
        clips, ns, n, hop = max(1, int(4096 * args.scale)), 80000, 512, 128
        w = hamming_periodic(n)
        xd, _ = device_batch(clips, ns, 20261017 + 8)
        nt = zaf.stft_geometry(ns, n, hop)[1]
        spec = zaf.empty((clips, nt, n), np.complex64)
        plan, _ = zaf._stft_plan(w, hop)
        lib, C = zaf._lib.lib(), zaf._lib.C
        cfg = f"{clips} clips x 5 s @ 16 kHz, N=512 hop=128"
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_stft_f32(
            plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 0, s.ptr)), args.steps)
        emit(line("stft-512", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * n * 8, nl))
        ylen = zaf.istft_geometry(n, nt, hop)[2]
        yd = zaf.empty((clips, ylen), np.float32)
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_istft_f32(
            plan, C.c_void_p(spec.ptr), clips, nt, 0, C.c_void_p(yd.ptr), ylen, s.ptr)), args.steps)
        emit(line("istft-512", cfg, clips * nt, "frames", ms, clips * nt * n * 8 + clips * ylen * 4, nl))
        xd.free(), spec.free(), yd.free()

    # ---- cfg 4: mdct + imdct, 2048 clips x 30 s @ 44.1 kHz, KBD N = 2048
    if only & {"mdct", "imdct"}:
        clips, ns, n = max(1, int(2048 * args.scale)), 1323000, 2048
        w = kbd(n)
        xd, _ = device_batch(clips, ns, 20261017 + 4)
        m, nt, _ = zaf.mdct_geometry(ns, n)
        plan, _ = zaf._mdct_plan(w)
        lib, C = zaf._lib.lib(), zaf._lib.C
        spec = zaf.empty((clips, nt, m), np.float32)
        cfg = f"cfg4: {clips} clips x 30 s @ 44.1 kHz, KBD N=2048"

        def f_mdct(stream):
            zaf._lib.check(lib.zafb_mdct_f32(plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 0, stream.ptr))

        ms, _, nl = timeit(f_mdct, args.steps)
        if "mdct" in only:
            emit(line("mdct", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * m * 4, nl))
        if "imdct" in only:
            ylen = zaf.imdct_geometry(m, nt)[1]
            pitch = (ylen + 1) & ~1
            yd = zaf.empty((clips, pitch), np.float32)

            def f_imdct(stream):
                zaf._lib.check(lib.zafb_imdct_f32(plan, C.c_void_p(spec.ptr), clips, nt, 0, C.c_void_p(yd.ptr), pitch, stream.ptr))

            ms, _, nl = timeit(f_imdct, args.steps)
            emit(line("imdct", cfg, clips * nt, "frames", ms, clips * nt * m * 4 + clips * ylen * 4, nl))
            yd.free()
        xd.free()
        spec.free()

    # ---- mdct + imdct at N = 1024 (the second window length with warp kernels)
    if "mdct1024" in only:
        clips, ns, n = max(1, int(2048 * args.scale)), 1323000, 1024
        w = kbd(n)
        xd, _ = device_batch(clips, ns, 20261017 + 9)
        m, nt, _ = zaf.mdct_geometry(ns, n)
        plan, _ = zaf._mdct_plan(w)
        lib, C = zaf._lib.lib(), zaf._lib.C
        spec = zaf.empty((clips, nt, m), np.float32)
        cfg = f"{clips} clips x 30 s @ 44.1 kHz, KBD N=1024"
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mdct_f32(
            plan, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(spec.ptr), 0, s.ptr)), args.steps)
        emit(line("mdct-1024", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * m * 4, nl))
        ylen = zaf.imdct_geometry(m, nt)[1]
        pitch = (ylen + 1) & ~1
        yd = zaf.empty((clips, pitch), np.float32)
        ms, _, nl = timeit(lambda s: zaf._lib.check(lib.zafb_imdct_f32(
            plan, C.c_void_p(spec.ptr), clips, nt, 0, C.c_void_p(yd.ptr), pitch, s.ptr)), args.steps)
        emit(line("imdct-1024", cfg, clips * nt, "frames", ms, clips * nt * m * 4 + clips * ylen * 4, nl))
        xd.free(), spec.free(), yd.free()

    # ---- cfg 3: melspectrogram + mfcc, 4096 clips x 5 s @ 16 kHz, N = 1024, hop = 256, 128 mels, 40 coefficients
    if only & {"mel", "mfcc", "meltc", "melf64"}:
        clips, ns, n, hop, fs = max(1, int(4096 * args.scale)), 80000, 1024, 256, 16000
        w = hamming_periodic(n)
        fb = zaf.melfilterbank(fs, n, 128)
        xd, _ = device_batch(clips, ns, 20261017 + 3)
        nt = zaf.stft_geometry(ns, n, hop)[1]
        cfg = f"cfg3: {clips} clips x 5 s @ 16 kHz, N=1024 hop=256, 128 mels, 40 coeffs"
        lib, C = zaf._lib.lib(), zaf._lib.C
        plan_mel, _, _ = zaf._mel_plan(w, hop, fb, 0)
        plan_mfcc, _, _ = zaf._mel_plan(w, hop, fb, 40)
        od = zaf.empty((clips, nt, 128), np.float32)
        if "mel" in only:
            ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_melspectrogram_f32(
                plan_mel, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), args.steps)
            emit(line("melspectrogram", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 128 * 4, nl,
                      "FP32/shared-memory bound, not HBM (SURVEY 8d)"))
        if "mfcc" in only:
            ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mfcc_f32(
                plan_mfcc, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), args.steps)
            emit(line("mfcc", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 40 * 4, nl,
                      "FP32/shared-memory bound, not HBM (SURVEY 8d)"))
        if "melf64" in only:  # the float64 route (accuracy option for purely tonal material)
            plan64, _, _ = zaf._mel_plan(w, hop, fb, 40, "fused", "float64")
            ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mfcc_f32(
                plan64, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), max(2, args.steps // 5))
            emit(line("mfcc-float64", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 40 * 4, nl,
                      "spectrum, filterbank sums and logarithm in FP64 (generic block kernel)"))
        if "meltc" in only:  # the dense tensor-core route of the same two transforms
            plan_mel_tc, _, _ = zaf._mel_plan(w, hop, fb, 0, "tensor")
            plan_mfcc_tc, _, _ = zaf._mel_plan(w, hop, fb, 40, "tensor")
            ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_melspectrogram_f32(
                plan_mel_tc, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), args.steps)
            emit(line("melspectrogram-tensor", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 128 * 4, nl,
                      "dense 3xTF32 tcgen05 filterbank product, spectra staged through L2"))
            ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mfcc_f32(
                plan_mfcc_tc, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), args.steps)
            emit(line("mfcc-tensor", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 40 * 4, nl,
                      "dense 3xTF32 tcgen05 filterbank + DCT products"))
        xd.free()

    # ---- the reference's own mel / mfcc example (zaf.py:347-357, 402-418): 44.1 kHz, N = 2048, hop = N/2, 128 mels, 20 coefficients
    if "mel2048" in only:
        clips, ns, n, hop, fs = max(1, int(1024 * args.scale)), 441000, 2048, 1024, 44100
        w = hamming_periodic(n)
        fb = zaf.melfilterbank(fs, n, 128)
        xd, _ = device_batch(clips, ns, 20261017 + 7)
        nt = zaf.stft_geometry(ns, n, hop)[1]
        cfg = f"reference example: {clips} clips x 10 s @ 44.1 kHz, N=2048 hop=1024, 128 mels, 20 coeffs"
        lib, C = zaf._lib.lib(), zaf._lib.C
        plan_mel, _, _ = zaf._mel_plan(w, hop, fb, 0)
        plan_mfcc, _, _ = zaf._mel_plan(w, hop, fb, 20)
        od = zaf.empty((clips, nt, 128), np.float32)
        ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_melspectrogram_f32(
            plan_mel, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), args.steps)
        emit(line("melspectrogram-2048", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 128 * 4, nl))
        ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_mfcc_f32(
            plan_mfcc, C.c_void_p(xd.ptr), clips, ns, ns, C.c_void_p(od.ptr), 0, s.ptr)), args.steps)
        emit(line("mfcc-2048", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * 20 * 4, nl))
        xd.free(), od.free()

    # ---- cfg 5: cqtspectrogram, 512 clips x 20 s @ 44.1 kHz, 12 bins/octave C1-C8, 25 frames/s
    if only & {"cqt", "cqttc"}:
        clips, ns, fs = max(1, int(512 * args.scale)), 882000, 44100
        kern = zaf.cqtkernel(fs, 12, 32.70319566257483, 4186.009044809578)
        xd, _ = device_batch(clips, ns, 20261017 + 5)
        step, nt, _, _ = zaf.cqt_geometry(ns, fs, 25, kern.shape[1])
        cfg = f"cfg5: {clips} clips x 20 s @ 44.1 kHz, 84 bins, L={kern.shape[1]}, step={step}"
        lib, C = zaf._lib.lib(), zaf._lib.C
        plan_cqt, nf, _ = zaf._cqt_plan(kern, step)
        od = zaf.empty((clips, nt, nf), np.float32)
        ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_cqt_f32(
            plan_cqt, C.c_void_p(xd.ptr), clips, ns, ns, 0, C.c_void_p(od.ptr), 0, s.ptr)), max(3, args.steps // 3))
        emit(line("cqtspectrogram", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * kern.shape[0] * 4, nl,
                  "FP32/shared-memory bound: 32768-point FFT per frame (SURVEY 8d)"))
        if "cqttc" in only:
            plan_tc, _, _ = zaf._cqt_plan(kern, step, "tensor")
            ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_cqt_f32(
                plan_tc, C.c_void_p(xd.ptr), clips, ns, ns, 0, C.c_void_p(od.ptr), 0, s.ptr)), max(2, args.steps // 5))
            emit(line("cqtspectrogram-tensor", cfg, clips * nt, "frames", ms, clips * ns * 4 + clips * nt * kern.shape[0] * 4, nl,
                      "kernel applied as a dense 3xTF32 tcgen05 product, spectra staged through HBM"))
        xd.free()

    # ---- dct / dst: 2^20 vectors of 1024 samples (the reference's example length)
    if "dct" in only:
        batch, n = max(1, int((1 << 20) * args.scale)), 1024
        xd, _ = device_batch(batch, n, 20261017 + 6, distinct=4096)
        lib, C = zaf._lib.lib(), zaf._lib.C
        od = zaf.empty((batch, n), np.float32)
        for kind, name in ((0, "dct"), (1, "dst")):
            for t in (1, 2, 3, 4):
                plan = zaf._dct_plans.get((kind, t, n), kind, t, n)
                ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_dct_f32(
                    plan, C.c_void_p(xd.ptr), batch, n, C.c_void_p(od.ptr), n, s.ptr)), max(3, args.steps // 3))
                emit(line(f"{name}-{t}", f"{batch} vectors x {n}", batch, "vectors", ms, batch * n * 8, nl))
        xd.free()

    # ---- dct / dst at the other warp-kernel lengths (types II-IV)
    if "dctsizes" in only:
        lib, C = zaf._lib.lib(), zaf._lib.C
        for n in (512, 2048, 4096):
            batch = max(1, int((1 << 30) // (4 * n) * args.scale))
            xd, _ = device_batch(batch, n, 20261017 + n, distinct=1024)
            od = zaf.empty((batch, n), np.float32)
            for kind, name in ((0, "dct"), (1, "dst")):
                for t in (2, 3, 4):
                    plan = zaf._dct_plans.get((kind, t, n), kind, t, n)
                    ms, o, nl = timeit(lambda s: zaf._lib.check(lib.zafb_dct_f32(
                        plan, C.c_void_p(xd.ptr), batch, n, C.c_void_p(od.ptr), n, s.ptr)), max(3, args.steps // 3))
                    emit(line(f"{name}-{t}-n{n}", f"{batch} vectors x {n}", batch, "vectors", ms, batch * n * 8, nl))
            xd.free(), od.free()

    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            for d in out_lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()

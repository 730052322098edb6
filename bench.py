#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: STFT frames/sec (win=2048 hop=512 fp32), % of HBM roofline.

A "step" is one forward STFT (the fused window+FFT kernel, zaf.stft) over one batch of
BASELINE cfg 2: 1024 clips x 10 s @ 48 kHz fp32 per GPU, Hamming window 2048, hop 512, the
full two-sided complex64 spectrum as the reference returns it.  Inputs are resident in HBM
when the timed region starts (`value`); `e2e` repeats the measurement through the public
drop-in call with pinned HOST buffers (H2D and D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU algorithm
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Multi-GPU: one process per GPU, clips are sharded by rank with no data-path collective (weak
scaling: every rank transforms its own 1024-clip batch); torch.distributed is used only for the
barrier and the max-over-ranks of the device-timed duration.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WIN, HOP, FS, SECONDS, CLIPS = 2048, 512, 48000, 10, 1024
NS = FS * SECONDS
SEED = 20261017 + 2
METRIC = "stft_frames_per_sec_win2048_hop512_fp32"


def hamming_periodic(n):
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)


# ----------------------------------------------------------------------------- CPU baseline
CPU_CLIPS_PER_CORE = 12  # ~0.9 core-seconds each: 16 cores -> ~14 core-seconds of CPU work per sample


def _cpu_worker(args):
    seed, clips = args
    import oracle  # the CPU port of the reference algorithm: checker / baseline only

    rng = np.random.default_rng(seed)
    w = hamming_periodic(N_WIN)
    xs = [rng.uniform(-1, 1, NS).astype(np.float32) for _ in range(clips)]  # inputs resident before timing
    frames = 0
    t0 = time.perf_counter()
    for x in xs:
        frames += oracle.stft(x, w, HOP).shape[1]
    return frames, time.perf_counter() - t0


def cpu_baseline(clips_per_core=CPU_CLIPS_PER_CORE, cores=None):
    """Time the oracle's port of zaf.stft (same operation sequence as zaf.py:95-141: Python framing
    loop + float64 pocketfft c2c) on all host cores over a bounded sample of the cfg-2 workload.
    Every process transforms its own clips; the duration is the slowest process's compute time."""
    cores = cores or len(os.sched_getaffinity(0))
    jobs = [(SEED + 1000 + i, clips_per_core) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, 1)] * cores)  # start the workers, import numpy, warm pocketfft
        res = pool.map(_cpu_worker, jobs)
    frames = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    return {
        "value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
        "sample": f"{cores * clips_per_core} clips x {SECONDS} s @ {FS} Hz of the {CLIPS}-clip batch "
                  f"({frames} frames, {dt:.2f} s wall = {sum(r[1] for r in res):.1f} core-seconds, {cores} processes, "
                  f"float64 NumPy pocketfft)",
    }, dt


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread
    (nvidia_ml_py, every ~2 ms), falling back to `nvidia-smi -lms` if NVML cannot be loaded."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            self.idx = int(vis.split(",")[gpu_index]) if vis else gpu_index
        except (ValueError, IndexError):
            self.idx = gpu_index
        self.samples = []  # (time, sm_mhz, reasons bitmask)
        self.smax = None
        self._stop = threading.Event()
        self._thread = None
        self.proc = None
        self.f = None

    def _poll(self, nv, handle):
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((time.time(), float(mhz), int(why)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self._thread.start()
            return
        except Exception:
            self._thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, t_begin, t_end):
        out = {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0, "source": None}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            inside = [s for s in self.samples if t_begin <= s[0] <= t_end]
            if not inside:  # a region shorter than one polling period: take the closest samples
                inside = sorted(self.samples, key=lambda s: abs(s[0] - 0.5 * (t_begin + t_end)))[:3]
            mask = 0
            for s in inside:
                mask |= s[2]
            if inside:
                out.update(sm_mhz=float(np.median([s[1] for s in inside])), samples=len(inside), source="nvml",
                           reasons=sorted(n for b, n in self.REASONS.items() if mask & b))
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime

        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            p = [s.strip() for s in line.split(",")]
            if len(p) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if not (t_begin - 0.05 <= ts <= t_end + 0.05):
                    continue
                sm.append(float(p[2]))
                smax.append(float(p[3]))
            except ValueError:
                continue
            for name, flag in zip(names, p[6:10]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), samples=len(sm), source="nvidia-smi")
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------- distributed plumbing
class Dist:
    def __init__(self, want):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.td = None
        if self.world > 1:
            import torch
            import torch.distributed as td

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            td.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.td, self.torch, self.backend = td, torch, backend
        if want != self.world and self.rank == 0 and self.world > 1:
            print(f"warning: --gpus {want} but WORLD_SIZE={self.world}", file=sys.stderr)

    def barrier(self):
        if self.td:
            self.td.barrier()

    def max(self, v):
        if not self.td:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64,
                              device="cuda" if self.backend == "nccl" else "cpu")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.td:
            self.td.barrier()
            self.td.destroy_process_group()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "stft_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ----------------------------------------------------------------------------- arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    cores = len(os.sched_getaffinity(0))
    per_core = max(1, args.cpu_clips_per_core)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(1, cores)
    vals, times = [], []
    for _ in range(args.steps):
        cb, dt = cpu_baseline(per_core, cores)
        vals.append(cb["value"])
        times.append(dt)
    value = float(np.mean(vals))
    cb["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE cfg 2: zaf.stft per clip, 10 s @ 48 kHz, Hamming 2048, hop 512 "
                               "(each step = a bounded sample of the 1024-clip batch)",
                   "window_length": N_WIN, "step_length": HOP, "clips_per_step": per_core * cores},
        "cpu_baseline": cb, "gpu_launches": 0,
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def split_merge_leg(zaf, dist, xd, w, nt, clips, stream, reps=3):
    """scatter (NCCL) -> local STFT -> gather (NCCL) of ONE cfg-2 batch held by rank 0; device-timed, max over ranks."""
    ids = [zaf.dist.make_unique_id() if dist.rank == 0 else None]
    dist.td.broadcast_object_list(ids, src=0)
    comm = zaf.dist.Communicator(dist.rank, dist.world, ids[0])
    lo, hi = comm.shard_range(clips)
    shard = zaf.empty((hi - lo, NS), np.float32)
    full = zaf.empty((clips, nt, N_WIN), np.complex64) if dist.rank == 0 else None
    spec_buf = zaf.empty((hi - lo, nt, N_WIN), np.complex64)
    e0, e1, e2, e3 = zaf.Event(), zaf.Event(), zaf.Event(), zaf.Event()
    best = None
    for _ in range(reps + 1):  # first pass = warm-up (NCCL channel set-up)
        dist.barrier()
        e0.record(stream)
        comm.scatter(xd if dist.rank == 0 else None, clips, (NS,), np.float32, stream=stream, out=shard)
        e1.record(stream)
        spec = zaf.stft(shard, w, HOP, stream=stream, out=spec_buf)
        e2.record(stream)
        comm.gather(spec, clips, stream=stream, out=full)
        e3.record(stream)
        e3.synchronize()
        t = [dist.max(e0.elapsed_ms(e1)), dist.max(e1.elapsed_ms(e2)), dist.max(e2.elapsed_ms(e3)), dist.max(e0.elapsed_ms(e3))]
        if best is None or t[3] < best[3]:
            best = t
    # the same merge without a collective: every rank's STFT kernel stores straight into rank 0's buffer
    # (CUDA IPC mapping, the stores cross NVLink), one barrier at the end
    direct = None
    try:
        view = comm.map_from_root(full, (clips, nt, N_WIN), np.complex64)
        mine = comm.rows(view, lo, hi)
        for _ in range(reps + 1):
            dist.barrier()
            e0.record(stream)
            comm.scatter(xd if dist.rank == 0 else None, clips, (NS,), np.float32, stream=stream, out=shard)
            e1.record(stream)
            zaf.stft(shard, w, HOP, stream=stream, out=mine)
            comm.barrier(stream)
            e3.record(stream)
            e3.synchronize()
            t = [dist.max(e0.elapsed_ms(e1)), dist.max(e1.elapsed_ms(e3)), dist.max(e0.elapsed_ms(e3))]
            if direct is None or t[2] < direct[2]:
                direct = t
        dist.barrier()
        comm.unmap(view)
    except Exception as exc:  # noqa: BLE001
        direct = f"{type(exc).__name__}: {exc}"[:200]
    comm.close()
    shard.free()
    spec_buf.free()
    if full is not None:
        full.free()
    return {"scaling": "strong", "global_clips": clips, "scatter_ms": best[0], "stft_ms": best[1], "gather_ms": best[2],
            "total_ms": best[3], "frames_per_sec": clips * nt / (best[3] * 1e-3),
            "scatter_bytes": int(clips * NS * 4 * (dist.world - 1) / dist.world),
            "gather_bytes": int(clips * nt * N_WIN * 8 * (dist.world - 1) / dist.world),
            "note": "rank 0 holds the batch; grouped ncclSend/ncclRecv over NVLink; best of %d" % reps,
            "direct_store_merge": ({"scatter_ms": direct[0], "stft_into_root_ms": direct[1], "total_ms": direct[2],
                                    "frames_per_sec": clips * nt / (direct[2] * 1e-3),
                                    "note": "no gather: each rank's STFT kernel writes into rank 0's HBM through a CUDA-IPC mapping"}
                                   if isinstance(direct, list) else {"error": direct})}


def run_ours(args):
    dist = Dist(args.gpus)
    cb = None
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline(max(1, args.cpu_clips_per_core))  # before any CUDA call (fork-safe)

    import zaf_python_b200 as zaf

    zaf.init(dist.local_rank)
    clips = args.clips
    w = hamming_periodic(N_WIN)
    nt = zaf.stft_geometry(NS, N_WIN, HOP)[1]
    frames = clips * nt
    rng = np.random.default_rng(SEED + dist.rank)
    pin_x = zaf.PinnedArray((clips, NS), np.float32)
    chunk = 64
    for c0 in range(0, clips, chunk):
        c1 = min(clips, c0 + chunk)
        pin_x.array[c0:c1] = rng.uniform(-1, 1, (c1 - c0, NS)).astype(np.float32)
    xd = zaf.to_device(pin_x.array)
    stream = zaf.Stream()
    out = zaf.empty((clips, nt, N_WIN), np.complex64)
    plan, _ = zaf._stft_plan(w, HOP)
    lib, C = zaf._lib.lib(), zaf._lib.C

    def step():
        zaf._lib.check(lib.zafb_stft_f32(plan, C.c_void_p(xd.ptr), clips, NS, NS, C.c_void_p(out.ptr),
                                         zaf.LAYOUT_FRAME_MAJOR, stream.ptr))

    sampler = ClockSampler(dist.local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    stream.synchronize()
    zaf.synchronize()
    dist.barrier()
    e0, e1 = zaf.Event(), zaf.Event()
    launches0 = zaf.launch_count()
    t_begin = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.synchronize()
    zaf.synchronize()
    t_end = time.time()
    launches = zaf.launch_count() - launches0
    local_ms = e0.elapsed_ms(e1)
    dist.barrier()
    total_ms = dist.max(local_ms)
    clocks = sampler.stop(t_begin, t_end)
    ms_per_step = total_ms / args.steps
    value = frames * dist.world / (ms_per_step * 1e-3)

    # correctness of what was just timed: two clips against the oracle (checker only)
    parity = None
    if dist.rank == 0:
        import oracle

        got = np.empty((nt, N_WIN), np.complex64)
        worst = 0.0
        for c in (0, clips - 1):
            zaf._lib.check(lib.zafb_memcpy_d2h(got.ctypes.data, C.c_void_p(out.ptr + c * nt * N_WIN * 8),
                                               got.nbytes, None))
            zaf.synchronize()
            worst = max(worst, *oracle.parity_metrics(got.T, oracle.stft(pin_x.array[c], w, HOP)))
        parity = worst
        assert worst <= 1e-5, f"parity broken: {worst}"

    # the same launch held for ~1 s: the board reaches its power cap and the SM clock drops; reported next to the headline
    extra = {}
    if args.sustained_steps > 0:
        sampler2 = ClockSampler(dist.local_rank)
        sampler2.start()
        tb = time.time()
        e0.record(stream)
        for _ in range(args.sustained_steps):
            step()
        e1.record(stream)
        e1.synchronize()
        te = time.time()
        sus_ms = dist.max(e0.elapsed_ms(e1)) / args.sustained_steps
        extra["sustained"] = {"steps": args.sustained_steps, "ms_per_step": sus_ms, "frames_per_sec": frames * dist.world / (sus_ms * 1e-3),
                              "clocks": sampler2.stop(tb, te)}

    # the other direction (istft) and the round trip, for the record
    yd = zaf.empty((clips, zaf.istft_geometry(N_WIN, nt, HOP)[2]), np.float32)

    def istep():
        zaf._lib.check(lib.zafb_istft_f32(plan, C.c_void_p(out.ptr), clips, nt, zaf.LAYOUT_FRAME_MAJOR,
                                          C.c_void_p(yd.ptr), yd.shape[1], stream.ptr))

    for _ in range(2):
        istep()
    stream.synchronize()
    k2 = max(3, args.steps // 4)
    e0.record(stream)
    for _ in range(k2):
        istep()
    e1.record(stream)
    e1.synchronize()
    extra["istft_ms_per_step"] = e0.elapsed_ms(e1) / k2
    extra["istft_frames_per_sec"] = frames / (extra["istft_ms_per_step"] * 1e-3)
    extra["round_trip_frames_per_sec"] = frames / ((ms_per_step + extra["istft_ms_per_step"]) * 1e-3)
    yd.free()

    # end to end through the public drop-in call, host buffers, H2D + D2H inside the timed region
    e2e = None
    e2e_clips = min(clips, args.e2e_clips)
    try:
        if args.e2e_steps <= 0:
            raise RuntimeError("skipped (--e2e-steps 0)")
        pin_out = zaf.PinnedArray((e2e_clips, nt, N_WIN), np.complex64)
        x_host = pin_x.array[:e2e_clips]
        zaf.stft(x_host, w, HOP, out=pin_out.array)
        dist.barrier()
        b0 = zaf.host_copy_bytes()
        t0 = time.perf_counter()
        ksteps = max(1, args.e2e_steps)
        for _ in range(ksteps):
            zaf.stft(x_host, w, HOP, out=pin_out.array)
        e2e_s = dist.max((time.perf_counter() - t0) / ksteps)
        b1 = zaf.host_copy_bytes()
        # bytes that crossed PCIe, counted by the library where it enqueues the copies.  The result in host memory is the
        # full two-sided spectrum (result_bytes); for large frame-major results only bins 0..N/2 are copied and host
        # threads write the Hermitian mirror (include/zafb200.h, zafb_stft_host_f32).
        e2e = {"value": e2e_clips * nt * dist.world / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": (b1[0] - b0[0]) // ksteps, "d2h_bytes_per_step": (b1[1] - b0[1]) // ksteps,
               "result_bytes_per_step": int(pin_out.nbytes),
               "clips_per_step": e2e_clips, "ms_per_step": 1e3 * e2e_s,
               "api": "zaf.stft(x_host, w, hop, out=pinned) -> zafb_stft_host_f32"}
        if dist.rank == 0:  # what the timed calls left in host memory, against the oracle (checker only)
            import oracle

            worst = 0.0
            for c in (0, e2e_clips - 1):
                worst = max(worst, *oracle.parity_metrics(pin_out.array[c].T, oracle.stft(x_host[c], w, HOP)))
            e2e["parity_max_rel_err"] = worst
            assert worst <= 1e-5, f"e2e parity broken: {worst}"
        # the same call returning the reference's own memory order (C-order (N, nt) per clip), reported beside it
        out_c = pin_out.array.reshape(e2e_clips, N_WIN, nt)
        zaf.stft(x_host, w, HOP, out=out_c, layout="bin_major")
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            zaf.stft(x_host, w, HOP, out=out_c, layout="bin_major")
        c_s = dist.max((time.perf_counter() - t0) / 2)
        e2e["c_order"] = {"value": e2e_clips * nt * dist.world / c_s, "ms_per_step": 1e3 * c_s, "layout": "bin_major"}
        if dist.rank == 0:
            e2e["c_order"]["parity_max_rel_err"] = max(oracle.parity_metrics(out_c[0], oracle.stft(x_host[0], w, HOP)))
            assert e2e["c_order"]["parity_max_rel_err"] <= 1e-5
        pin_out.free()
    except (MemoryError, RuntimeError) as exc:  # e.g. not enough pinnable host memory
        e2e = {"value": None, "unit": "frames/s", "error": str(exc)[:200]}

    # N > 1: the batch split / merge leg (strong scaling): rank 0 holds the whole cfg-2 batch in HBM, scatters the
    # clips over NCCL, every rank transforms its shard, the spectra are gathered back on rank 0
    if dist.world > 1 and not args.no_split_merge:
        try:
            extra["split_merge"] = split_merge_leg(zaf, dist, xd, w, nt, clips, stream)
        except Exception as exc:  # noqa: BLE001 -- the headline number does not depend on this leg
            extra["split_merge"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if dist.rank == 0:
        algo_bytes = clips * NS * 4 + frames * N_WIN * 8
        peak, peak_src = measured_peak()
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": dist.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE cfg 2 forward STFT: {clips} clips x {SECONDS} s @ {FS} Hz fp32 per GPU, "
                                   f"Hamming window {N_WIN}, hop {HOP}, full two-sided complex64 spectrum "
                                   f"({nt} frames/clip)",
                       "window_length": N_WIN, "step_length": HOP, "clips_per_gpu": clips, "global_clips": clips * dist.world,
                       "parallelism": f"clip-sharded x{dist.world}, no data-path collective",
                       "layout": "frame_major", "l2": "inputs+outputs per step (17.7 GB) exceed L2 (126 MB); no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel": "stft_warp_kernel<2048, false, 6>"},
            "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_max_rel_err": parity, "extra": extra,
        }
        print(json.dumps(line))
    dist.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 50; 5 for --impl reference)")
    ap.add_argument("--sustained-steps", type=int, default=300,
                    help="extra back-to-back launches (~1 s) reported as extra.sustained; 0 to skip")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS, help="clips per GPU (BASELINE cfg 2: 1024)")
    ap.add_argument("--e2e-clips", type=int, default=CLIPS)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-split-merge", action="store_true", help="N > 1: skip the NCCL split/merge leg")
    ap.add_argument("--cpu-clips-per-core", type=int, default=CPU_CLIPS_PER_CORE,
                    help="size of the bounded CPU sample (clips per host core)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 50
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

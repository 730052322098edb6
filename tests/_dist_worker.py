"""Worker for tests/test_dist.py: one process per GPU, launched with RANK / WORLD_SIZE / LOCAL_RANK /
MASTER_ADDR / MASTER_PORT in the environment.  Splits a batch from rank 0 over NCCL, transforms the
local shard, merges the spectra on rank 0 and compares them with the unsharded run bit for bit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zaf_python_b200 as zaf  # noqa: E402


def main():
    out_path = sys.argv[1]
    comm = zaf.dist.Communicator.from_env()
    rank, world = comm.rank, comm.world
    clips, ns, n, hop = 11, 24000, 2048, 512          # ragged shards on purpose
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, (clips, ns)).astype(np.float32)
    xd = zaf.to_device(x) if rank == 0 else None
    shard = comm.scatter(xd, clips, (ns,), np.float32)
    lo, hi = comm.shard_range(clips)
    assert shard.shape == (hi - lo, ns)
    spec = zaf.stft(shard, w, hop)
    full = comm.gather(spec, clips)
    every = comm.allgather(spec, clips)
    # merge without a collective: every rank's kernel stores straight into the root's buffer (CUDA IPC over NVLink)
    nt = spec.mem_shape[1]
    direct = zaf.empty((clips, nt, n), np.complex64) if rank == 0 else None
    view = comm.map_from_root(direct, (clips, nt, n), np.complex64)
    zaf.stft(shard, w, hop, out=comm.rows(view, lo, hi))
    zaf.synchronize()
    comm.barrier()
    # one-sided merges: (a) NCCL gather of bins 0 .. N/2 + device mirror kernel on the root, (b) every peer's one-sided
    # kernel stores the lower half of its frames straight into the root's two-sided buffer, the root mirrors in place
    half = zaf.stft(shard, w, hop, onesided=True)
    merged_half = comm.gather_onesided(half, clips, n)
    lib, C = zaf._lib.lib(), zaf._lib.C
    plan, _ = zaf._stft_plan(w, hop)
    direct2 = zaf.empty((clips, nt, n), np.complex64) if rank == 0 else None
    view2 = comm.map_from_root(direct2, (clips, nt, n), np.complex64)
    mine2 = comm.rows(view2, lo, hi)
    if rank == 0:
        zaf.stft(shard, w, hop, out=mine2)
    else:
        zaf._lib.check(lib.zafb_stft_onesided_f32(plan, C.c_void_p(shard.ptr), hi - lo, ns, shard.pitch, C.c_void_p(mine2.ptr), n, None))
    zaf.synchronize()
    comm.barrier()
    if rank == 0 and hi < clips:
        tail = comm.rows(direct2, hi, clips)
        zaf._lib.check(lib.zafb_spec_mirror_f32(C.c_void_p(tail.ptr), n, (clips - hi) * nt, n, C.c_void_p(tail.ptr), None))
        zaf.synchronize()
    table = zaf.to_device(np.arange(1000, dtype=np.float32) * (1.0 if rank == 0 else 0.0))
    comm.broadcast(table)
    slowest = comm.max(float(rank + 1))
    zaf.synchronize()
    res = {"rank": rank, "world": world, "max": slowest, "table_ok": bool(np.array_equal(table.to_host(), np.arange(1000, dtype=np.float32)))}
    whole = zaf.stft(zaf.to_device(x), w, hop).to_host()
    res["allgather_bitwise"] = bool(np.array_equal(every.to_host(), whole))
    if rank == 0:
        res["gather_bitwise"] = bool(np.array_equal(full.to_host(), whole))
        res["shape"] = list(full.shape)
        res["direct_bitwise"] = bool(np.array_equal(np.swapaxes(direct.to_host(), 1, 2), whole))
        res["half_gather_bitwise"] = bool(np.array_equal(merged_half.to_host(), whole))
        res["half_direct_bitwise"] = bool(np.array_equal(np.swapaxes(direct2.to_host(), 1, 2), whole))
        res["nccl"] = zaf.dist.nccl_version()
    with open(f"{out_path}.{rank}", "w") as f:
        json.dump(res, f)
    comm.barrier()
    comm.unmap(view)
    comm.unmap(view2)
    comm.close()


if __name__ == "__main__":
    main()

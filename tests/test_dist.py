"""Batch split / merge (zaf-python_b200/_dist.py, csrc/dist.cu): shard arithmetic and the id
rendezvous on CPU, the NCCL collectives on the GPU (single rank always; two ranks when the box has
two GPUs)."""
import json
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_matches_c_abi():
    import ctypes as C

    import zaf_python_b200 as zaf

    lib = zaf._lib.lib()
    for n in (0, 1, 7, 11, 512, 1024, 4097):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                b, e = C.c_int64(-1), C.c_int64(-1)
                zaf._lib.check(lib.zafb_dist_shard_range(n, rank, world, C.byref(b), C.byref(e)))
                assert (b.value, e.value) == zaf.shard_range(n, rank, world)
    with pytest.raises(ValueError):
        zaf._lib.check(lib.zafb_dist_shard_range(4, 2, 2, None, None))


def test_id_rendezvous_three_ranks():
    from zaf_python_b200 import dist

    port = _free_port()
    payload = bytes(range(128))
    got = {}

    def run(rank):
        got[rank] = dist.exchange_id(rank, 3, lambda: payload, "127.0.0.1", port, timeout=30)

    threads = [threading.Thread(target=run, args=(r,)) for r in (2, 1, 0)]  # peers first: they retry until rank 0 listens
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert got == {0: payload, 1: payload, 2: payload}


@pytest.mark.gpu
def test_single_rank_collectives(zaf_gpu):
    zaf = zaf_gpu
    comm = zaf.dist.Communicator(0, 1, zaf.dist.make_unique_id())
    assert zaf.dist.nccl_version() >= 22000
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (5, 4096)).astype(np.float32)
    xd = zaf.to_device(x)
    shard = comm.scatter(xd, 5, (4096,), np.float32)
    assert np.array_equal(shard.to_host(), x)
    merged = comm.gather(shard, 5)
    assert np.array_equal(merged.to_host(), x)
    assert np.array_equal(comm.allgather(shard, 5).to_host(), x)
    assert np.array_equal(comm.broadcast(xd).to_host(), x)
    assert comm.max(2.5) == 2.5
    # one-sided merge on a single rank: gather = copy, then the device mirror kernel rebuilds the two-sided spectrum
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(1024) / 1024)
    half = zaf.stft(xd, w, 256, onesided=True)
    assert np.array_equal(comm.gather_onesided(half, 5, 1024).to_host(), zaf.stft(x, w, 256))
    comm.close()


@pytest.mark.gpu
def test_two_rank_split_merge_bitwise(zaf_gpu, tmp_path):
    if zaf_gpu.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    out = str(tmp_path / "res")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), out], cwd=ROOT,
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
    res = [json.load(open(f"{out}.{r}")) for r in range(2)]
    assert all(r["max"] == 2.0 and r["table_ok"] and r["allgather_bitwise"] for r in res), res
    assert res[0]["gather_bitwise"] and res[0]["direct_bitwise"] and res[0]["shape"] == [11, 2048, 48], res
    assert res[0]["half_gather_bitwise"] and res[0]["half_direct_bitwise"], res

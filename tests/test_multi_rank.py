"""N > 1 host logic on CPU: world_size-2 gloo run of the bench plumbing (barrier, max-over-ranks)
and of the clip sharding -- sharded results must equal the unsharded ones bit for bit."""
import hashlib
import json
import os
import socket
import subprocess
import sys

import numpy as np

import oracle
from zaf_python_b200 import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 512, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(1024, r, 8) for r in range(8)] == [(128 * r, 128 * (r + 1)) for r in range(8)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo(tmp_path):
    out = tmp_path / "ranks.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_rank_worker.py"), str(out)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    proc = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["world"] == 2 and res["max"] == 2.0
    parts = sorted(res["gathered"], key=lambda d: d["rank"])
    assert [tuple(p["range"]) for p in parts] == [(0, 3), (3, 7)]
    rng = np.random.default_rng(123)
    x = rng.uniform(-1, 1, (7, 3000)).astype(np.float32)
    w = oracle.hamming_periodic(256)
    whole = [hashlib.sha256(oracle.stft(x[c], w, 64).tobytes()).hexdigest() for c in range(7)]
    assert parts[0]["digests"] + parts[1]["digests"] == whole

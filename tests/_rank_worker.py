"""Worker for tests/test_multi_rank.py: launched by torch.distributed.run with the gloo backend."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import oracle  # noqa: E402
from zaf_python_b200 import shard_range  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist = bench.Dist(int(os.environ["WORLD_SIZE"]))
    assert dist.backend == "gloo"
    clips = 7
    lo, hi = shard_range(clips, dist.rank, dist.world)
    rng = np.random.default_rng(123)
    x = rng.uniform(-1, 1, (clips, 3000)).astype(np.float32)   # same batch on every rank
    w = oracle.hamming_periodic(256)
    digests = [hashlib.sha256(oracle.stft(x[c], w, 64).tobytes()).hexdigest() for c in range(lo, hi)]
    gathered = [None] * dist.world
    dist.td.all_gather_object(gathered, {"rank": dist.rank, "range": (lo, hi), "digests": digests})
    dist.barrier()
    slowest = dist.max(float(dist.rank + 1))
    if dist.rank == 0:
        with open(out_path, "w") as f:
            json.dump({"gathered": gathered, "max": slowest, "world": dist.world}, f)
    dist.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors for the window lengths added after the first set (warp kernels for N = 256, 512, 4096), from the
UNMODIFIED reference, in separate files so that the first set stays byte-identical:

    python tests/golden/make_golden_sizes.py      (build container only: needs /root/reference)

writes golden_stft_sizes.npz and golden_mdct_sizes.npz; tests/conftest.py merges `golden_<name>_sizes.npz` into
`golden_<name>.npz`, so every golden test (oracle on the CPU, kernels on the GPU) covers them.
"""
import os

import numpy as np

from make_golden import HERE, f32, hamming_p, load_reference


def kbd(n, alpha=5.0):
    k = np.kaiser(n // 2 + 1, np.pi * alpha)
    half = np.sqrt(np.cumsum(k[: n // 2]) / np.sum(k))
    return f32(np.concatenate([half, half[::-1]]))


def main():
    zaf = load_reference()
    rng = np.random.default_rng(20261017 + 4096)
    g = {}
    for name, (ns, n, hop) in {
        "rand_n4096_h1024": (6000, 4096, 1024),
        "rand_n4096_h2048": (9001, 4096, 2048),
        "rand_n512_h128": (3000, 512, 128),
        "rand_n256_h128": (1500, 256, 128),
    }.items():
        x, w = f32(rng.uniform(-1, 1, ns)), hamming_p(n)
        spec = zaf.stft(x, w, hop)
        g[f"{name}/x"], g[f"{name}/w"], g[f"{name}/hop"] = x, w, np.int64(hop)
        g[f"{name}/stft"] = spec
        g[f"{name}/istft"] = zaf.istft(spec, w, hop)
    m = {}
    for name, (ns, w) in {
        "rand_kbd4096": (10000, kbd(4096)),
        "rand_sine512": (3000, f32(np.sin(np.pi / 512 * (np.arange(512) + 0.5)))),
        "rand_kbd512": (2049, kbd(512)),
    }.items():
        x = f32(rng.uniform(-1, 1, ns))
        c = zaf.mdct(x, w)
        m[f"{name}/x"], m[f"{name}/w"] = x, w
        m[f"{name}/mdct"] = c
        m[f"{name}/imdct"] = zaf.imdct(c, w)
    for fname, arrays in (("stft_sizes", g), ("mdct_sizes", m)):
        path = os.path.join(HERE, f"golden_{fname}.npz")
        np.savez_compressed(path, **arrays)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays")


if __name__ == "__main__":
    main()

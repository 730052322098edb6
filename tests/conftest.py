import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


class Golden:
    """Access to tests/golden/golden_<name>.npz as nested 'case/field' keys."""

    def __init__(self, name):
        self._z = dict(np.load(os.path.join(GOLDEN_DIR, f"golden_{name}.npz")))
        extra = os.path.join(GOLDEN_DIR, f"golden_{name}_sizes.npz")  # later window lengths (make_golden_sizes.py)
        if os.path.exists(extra):
            self._z.update(np.load(extra))

    def cases(self):
        return sorted({k.split("/")[0] for k in self._z if "/" in k})

    def get(self, case, field):
        v = self._z[f"{case}/{field}"]
        return v.item() if v.ndim == 0 else v

    def has(self, case, field):
        return f"{case}/{field}" in self._z

    def raw(self, key):
        return self._z[key]


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def _get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return _get


def _gpu_available():
    try:
        import zaf_python_b200 as zaf

        return zaf.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def zaf_gpu():
    """The product package, initialised on device 0.  GPU tests fail loudly if the native
    library is missing -- there is no fallback."""
    import zaf_python_b200 as zaf

    zaf._lib.lib()  # a missing libzafb200.so is an error, never a skip
    if not _gpu_available():
        pytest.skip("libzafb200.so is built but this machine has no CUDA device (run with -m 'not gpu')")
    zaf.init(0)
    return zaf

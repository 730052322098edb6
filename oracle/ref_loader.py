"""Import the vendored, unmodified reference module ``oracle/_ref/zaf.py`` (see ``oracle/make_ref.py``).

TEST INFRASTRUCTURE ONLY.  The reference's single missing import in this image is ``matplotlib`` (zaf.py:42), which the
transform hot path never touches; an empty stand-in module is registered before the import.  The reference's examples
call ``scipy.signal.hamming`` (removed from SciPy >= 1.13) -- callers here build their windows themselves.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "zaf.py")
_cached = None


def available() -> bool:
    return os.path.exists(REF_PATH)


def load():
    """The reference module, or None when ``oracle/_ref/zaf.py`` has not been built."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    manifest = os.path.join(_HERE, "_ref", "MANIFEST.json")
    if os.path.exists(manifest):  # the copy must still be the file the recipe hashed
        with open(manifest) as f, open(REF_PATH, "rb") as g:
            if json.load(f).get("sha256") != hashlib.sha256(g.read()).hexdigest():
                raise RuntimeError("oracle/_ref/zaf.py does not match its manifest: rebuild it with oracle/make_ref.py")
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("zaf_reference", REF_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod

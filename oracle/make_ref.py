#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference module, vendored at build time.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference is one pure-Python file (``/root/reference/zaf.py``), so "building the reference" is a byte-for-byte
copy of that file to ``oracle/_ref/zaf.py`` plus a manifest with its SHA-256.  ``oracle/_ref/`` is git-ignored (the
reference's source never enters this repository's history) but travels to the GPU box with the snapshot, like the
built ``.so`` -- there ``bench.py --impl reference`` / ``cpu_baseline`` time the reference ITSELF
(``cpu_baseline.kind == "reference"``) instead of the port in ``oracle/zaf_oracle.py``, and
``tests/test_oracle_golden.py::test_vendored_reference_*`` checks the port against it.

    python oracle/make_ref.py            # run in the build container (needs /root/reference)

``__graft_entry__.build()`` runs this whenever ``/root/reference`` exists.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ZAFB_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref(verbose=True):
    """Copy the reference module; returns the destination path, or None when /root/reference is absent."""
    src = os.path.join(REF_SRC, "zaf.py")
    if not os.path.exists(src):
        return None
    os.makedirs(DST, exist_ok=True)
    dst = os.path.join(DST, "zaf.py")
    with open(src, "rb") as f:
        blob = f.read()
    digest = hashlib.sha256(blob).hexdigest()
    manifest = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(dst) and os.path.exists(manifest):
        try:
            with open(manifest) as f:
                if json.load(f).get("sha256") == digest:
                    return dst
        except (OSError, ValueError):
            pass
    shutil.copyfile(src, dst)
    with open(manifest, "w") as f:
        json.dump({"source": src, "sha256": digest, "bytes": len(blob),
                   "note": "unmodified copy of the reference module; git-ignored, never edited"}, f, indent=1)
    if verbose:
        print(f"oracle/_ref/zaf.py <- {src} (sha256 {digest[:16]}...)")
    return dst


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)

"""CPU oracle for the zaf.py transform hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (the
``zaf-python_b200`` package and ``libzafb200.so``) imports, links or executes
anything in this directory.  The only allowed users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / CPU baseline.

Parity pin: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the unmodified reference
``/root/reference/zaf.py`` generated in the build container by
``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``
(see ``tests/test_oracle_golden.py``).
"""
from .zaf_oracle import *  # noqa: F401,F403

"""CPU oracle for the SS2D scan path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product package ``xfmamba_b200`` never does, and has no
CPU fallback.  See ``oracle/xfscan_oracle.c`` for the reference file:line each function restates and
``tests/test_oracle_golden.py`` for how the oracle is pinned against vectors produced by the
reference's own Python (``tests/golden/make_golden.py``).
"""
from .oracle import *  # noqa: F401,F403

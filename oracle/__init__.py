"""CPU oracle for the GGAD hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (numpy / scipy / torch-CPU) restatement of the reference
algorithm for the message-passing + outlier-synthesis path.  It exists so that
the CUDA path in ``ggad_b200`` can be checked; it is never the thing shipped or
measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``ggad_b200/`` imports it, and the product path raises if the CUDA library is
missing instead of falling back here.

Parity status: the reference ships no tests, golden vectors or fixtures for
this path, so the oracle is pinned against outputs of the
*reference modules themselves* (``/root/reference/model.py`` and
``/root/reference/src/graphsage.py``) imported in the build container and run on
small seeded graphs; the vectors are committed under ``tests/golden/`` together
with ``tests/golden/make_golden.py`` that generated them.
"""
from .ggad_oracle import *  # noqa: F401,F403

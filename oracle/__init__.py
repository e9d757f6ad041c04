"""CPU oracle for the multitaper -> CSM -> coherence / PLI / Granger hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``spectral_connectivity_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs do.
"""
from . import oracle  # noqa: F401

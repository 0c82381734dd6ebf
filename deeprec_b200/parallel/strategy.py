"""CollectiveStrategy facade (python/distribute/group_embedding_collective_strategy.py:28-138).
Filled in by parallel/collective.py; this module only holds the process-wide current strategy."""
from __future__ import annotations

_CURRENT = None


def current():
    return _CURRENT


def set_current(s) -> None:
    global _CURRENT
    _CURRENT = s

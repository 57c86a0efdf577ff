"""TEST INFRASTRUCTURE ONLY.

CPU oracles for the pattern-matching (MCC) hot path of nansencenter/sea_ice_drift.
Nothing under ``sea_ice_drift_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs do, and there only as the checker or the timed CPU baseline.
"""

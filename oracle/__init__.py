"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DIGAT dual-graph encoder hot path.

Nothing under ``digat_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
"""

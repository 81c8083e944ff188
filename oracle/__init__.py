"""CPU oracle for the solve/rotate/rule_n hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or the timed CPU
baseline.  The product package ``xmca_b200`` never imports this module.
"""

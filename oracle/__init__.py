"""CPU oracle for the RBniCS greedy-sweep hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product path (``rbnics_b200``) never does.
"""

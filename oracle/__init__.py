"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the MAFED distillation path.

Nothing under ``mafed_b200/`` may import this package.  The only permitted
callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or as
the timed CPU baseline, never as the thing shipped.
"""

"""CPU oracle for the Sepformer hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; the product package never does (tests/test_boundary.py checks).
"""

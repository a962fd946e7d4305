"""CPU oracle = test infrastructure.  Restates the reference's algorithm for the hot path
(embedding gather -> Koopman forecast -> MPC solve -> portfolio step -> metrics) in numpy / fp64.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import anything from here; the product package never does (tests/test_boundary.py checks)."""

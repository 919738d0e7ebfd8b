"""CPU oracle for the edge-enhancement + PGD-step hot path.

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and bench.py's
``cpu_baseline`` / ``--impl reference`` legs.  The product package (edge_enhancement_b200/)
must never import this package; it has no CPU fallback.

Parity pin: the reference repo has no tests or golden vectors; the oracle is
pinned against the reference's own modules executed in the build container
(oracle/make_golden.py -> tests/golden/*.npz).
"""

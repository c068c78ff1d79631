"""CPU oracle for the pySOT RBF-surrogate hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pysot_b200/`` may import this
package.  The only allowed users are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only
as the checker or the CPU baseline, never as the thing shipped.

Parity status: PINNED.  ``tools/make_golden.py`` imports the unmodified
reference from ``/root/reference`` in the build container, runs it and this
oracle on the same seeded inputs, asserts agreement, and stores the reference's
outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the
oracle against those fixtures wherever the suite runs.
"""

"""CPU oracle for the rlpyt hot path (TEST INFRASTRUCTURE - never shipped).

A numpy / torch-CPU restatement of the reference algorithms on the path named by
BASELINE.json:north_star.  Every function cites the reference file:line it follows
(paths relative to the reference checkout, commit f04f23d).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` - and there only as the
checker or as the CPU baseline being timed, never as the thing shipped.  Nothing
under ``rlpyt_b200/`` imports it; the product path raises if the CUDA library is
missing instead of falling back to this code.

Pinning: the reference's own test-suite holds no numerical fixture for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified
reference imported in the build container: ``tests/golden/make_golden.py`` generates
``tests/golden/*.npz`` from ``/root/reference`` and ``tests/test_oracle_golden.py``
checks every oracle function against them (bit-exact), plus the hand-checkable
known-answer vectors of SURVEY.md section 9.
"""

"""CPU oracle for the stacked-hourglass hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hourglass-pose-estimation_b200/``
(the product) may import this package; it is used by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs as the checker and CPU baseline.

Every function is a restatement (not a copy) of the reference algorithm and
cites the reference file:line it follows.  Parity is pinned by
``tests/golden/*.npz`` -- outputs of the LIVE reference (imported from
/root/reference in the dev container by ``oracle/make_golden.py``) on seeded
inputs; ``tests/test_oracle_golden.py`` checks the oracle against them.
The reference ships no tests / golden vectors of its own.
"""

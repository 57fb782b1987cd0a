"""CPU oracle for the SubGAcc hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing in
``surel_plus_b200`` does, and the product path has no CPU fallback.

Parity status: PINNED -- ``oracle/subg_oracle.c`` and ``oracle/pyoracle.py`` are
checked against the unmodified reference compiled/imported from
``/root/reference`` (tests/test_oracle_vs_reference.py, runs where the reference
tree exists) and against the committed fixtures ``tests/golden/*.npz`` that
``tests/golden/make_golden.py`` generated from the reference.
"""

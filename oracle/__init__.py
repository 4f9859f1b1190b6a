"""CPU oracle for the EfficientAT hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain functional PyTorch (CPU, fp32 or fp64), the
algorithm of the reference's mel front end and MN / DyMN networks.  It exists to
check the CUDA product path; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``efficientat_b200/`` imports it, and the product path raises if
its CUDA library is missing -- there is no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference modules themselves,
imported from /root/reference in the build container by
``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.
"""

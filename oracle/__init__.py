"""CPU oracle for the DiffPIR sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline, never as the thing shipped.  The product path
(``diffpir_b200``) must not import this package.

Each function restates one piece of the reference algorithm (plain torch on
CPU, fp32 or fp64) and cites the reference file:line it follows.  Parity
pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference's own modules,
imported unmodified from /root/reference in the build container by
``oracle/gen_golden.py`` (fixtures committed under ``tests/golden/``) and
re-checked live by ``tests/test_oracle_vs_reference.py`` whenever
/root/reference is present.
"""

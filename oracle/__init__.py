"""CPU oracle for the CompeteSMoE sparse-MoE layer hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain-PyTorch (CPU) restatement of the reference's algorithm for the path
named in BASELINE.json (`north_star`); every function cites the reference file:line it follows.  It may be imported
only by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s `cpu_baseline` / `--impl reference` legs, and there
only as the checker or the timed CPU baseline -- never by `competesmoe_b200/` (the product), which has no CPU path.

Parity pinning: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against outputs of the reference itself: `oracle/gen_golden.py` imports the unmodified reference modules from
/root/reference in the build container, runs them on seeded inputs and stores inputs/outputs/gradients under
`tests/golden/`; `tests/test_oracle_golden.py` replays the oracle against those fixtures (CPU, no reference needed).
"""

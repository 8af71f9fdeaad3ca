"""CPU oracle for the stiffness-assembly + linear-solve path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy/scipy, the algorithm the reference (JWock82/Pynite,
`/root/reference/Pynite`) runs for `FEModel3D.Ke -> tocsr -> _partition -> spsolve -> reactions`.
It exists so the CUDA path can be checked on a GPU box where `/root/reference` is absent.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; the product package `pynite_b200` never does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against fixtures
under `tests/golden/` that were produced by importing and running the reference itself
(`tests/golden/make_golden.py`), and against the textbook values the reference's own tests hold.
"""

"""CPU oracle for the GoGP hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference's algorithm
(/root/reference/gp/gp.go, kernel/kernel.go, kernel/noise.go, gp/model.go and
the tutorial kernels).  It exists to check the CUDA path; nothing under
``gogp_b200/`` may import it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it.

Parity status: the Go reference cannot be built in this environment (no Go
toolchain, gonum v0.9.3 and infergo v1.2.2 sources absent), so the oracle is
pinned against the only known-answer vectors the reference holds for this
path -- the 7 ``TestProduce`` and 5 ``TestElementalModel`` cases of
gp/gp_test.go (6 printed digits, N <= 2, Normal kernel) -- and, beyond those,
against 50-digit mpmath evaluations of the same formulas
(tests/test_oracle_mpmath.py).  Periodic / Matern / composed kernels, N > 2 and
any digit past the sixth are "parity unpinned" by the reference itself.
"""

"""CPU oracle for SUPIR's EDM sampling hot path — TEST INFRASTRUCTURE, not product code.

A plain-PyTorch fp32 restatement of the reference algorithm (Fanghua-Yu/SUPIR; every function cites the reference
file:line it follows). It exists so that parity tests can run on the GPU box, where /root/reference does not exist.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against outputs of the
UNMODIFIED reference imported in the build container (tests/golden/ref_stubs.py): tests/golden/make_golden.py commits
small fixtures (seeded tiny-config weights + inputs + reference outputs) under tests/golden/, and
tests/test_oracle_vs_golden.py checks the oracle against them on every run; tests/test_oracle_vs_reference.py
re-derives them live whenever /root/reference is present.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (supir_b200/) never does: it fails loudly when its CUDA library is missing.
"""

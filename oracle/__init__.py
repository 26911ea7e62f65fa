"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's hybrid window-scoring path
(Ogunleyemma1/Hybrid-VAE-CNN-for-SHM).  Nothing in the product package
(`hybrid-vae-cnn-for-shm_b200/`) imports this; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may.  It is the checker, never the thing shipped or measured as the product.

Parity pin: the reference repo holds NO tests and NO known-answer vectors for
this path, and the arithmetic lives in PyTorch (no version
pinned by the reference).  The oracle is therefore pinned against outputs of the
reference's own `Models/` modules executed in the build container
(torch 2.11.0 CPU) on seeded inputs/weights/eps: `tests/golden/make_golden.py`
imports `/root/reference/*/Models/*.py`, runs them, and commits the vectors
under `tests/golden/`; `tests/test_oracle_golden.py` checks this restatement
against every one of those vectors.
"""

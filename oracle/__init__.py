"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's denoising hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package; nothing under `seervideoldm_b200/` does (tests/test_no_oracle_in_product.py checks).
"""

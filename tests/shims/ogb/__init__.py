"""Stand-in for the `ogb` package the reference imports (backend_pim/spmv.py:8-10, spmm_test.py:50, inference.py:56)."""

"""CPU oracle -- test infrastructure only (see oracle/numpy_oracle.py header)."""

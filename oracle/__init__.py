"""CPU oracle — test infrastructure only (see u1_oracle.py header)."""

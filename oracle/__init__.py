"""CPU oracle package -- test infrastructure only (see mt3_oracle.py header)."""

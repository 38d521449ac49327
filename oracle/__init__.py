"""CPU oracle package -- test infrastructure only (see attention_oracle.py)."""

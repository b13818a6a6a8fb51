"""CPU oracle of the hot path -- TEST INFRASTRUCTURE ONLY (see tcr_oracle.c / tcr_oracle.py headers)."""

"""Test infrastructure only.  See oracle/imm_oracle.py header."""

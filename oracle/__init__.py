"""CPU oracle for the pyrayt_b200 parity tests (test infrastructure only)."""

"""CPU oracle (test infrastructure only; never imported by ladiff_b200)."""

"""CPU oracle (test infrastructure only — see oracle/ao_oracle.c header)."""

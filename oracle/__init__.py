"""CPU oracle (test infrastructure only; see oracle/ba_oracle.c)."""

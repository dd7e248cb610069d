"""CPU oracle (test infrastructure only). See gamma_oracle.c for the contract."""

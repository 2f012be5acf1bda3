"""Parity oracle (test infrastructure only) -- see oracle/brotli_oracle.c."""

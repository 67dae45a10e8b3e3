"""CPU oracle (test infrastructure only).  see nimrud_oracle.py and oracle.c."""

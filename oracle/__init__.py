"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see fm_oracle.h).  PARITY UNPINNED: no runnable
reference, no reference golden vectors; checked against builder-authored exact KATs."""

"""CPU oracle for the barcode-match path -- test infrastructure only (see nr_oracle.c)."""

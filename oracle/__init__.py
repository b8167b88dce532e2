"""CPU oracle for the filtered k-NN hot path.  TEST INFRASTRUCTURE ONLY -- see oracle/hvs_oracle.c."""

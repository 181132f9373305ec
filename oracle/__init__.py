"""CPU oracle for the ark-mpc hot path — TEST INFRASTRUCTURE ONLY (see pyoracle.py / ark_oracle.c)."""

"""CPU oracle of the per-ray pipeline — TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header)."""

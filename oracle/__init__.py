"""CPU oracle for the UMHS hot path -- TEST INFRASTRUCTURE ONLY (never imported by the product package)."""

"""TEST INFRASTRUCTURE ONLY - the CPU oracle.  Not part of the product path."""

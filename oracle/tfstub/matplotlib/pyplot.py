"""matplotlib.pyplot placeholder (never called on the hot path).  TEST INFRASTRUCTURE ONLY."""

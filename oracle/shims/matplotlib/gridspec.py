"""Empty import stub (test infrastructure)."""

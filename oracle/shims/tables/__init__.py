"""Empty import stub (test infrastructure): the reference imports this at module load but the oracle runner bypasses all file I/O."""

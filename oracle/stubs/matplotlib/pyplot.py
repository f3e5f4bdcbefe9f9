"""Empty stand-in for matplotlib.pyplot (test infrastructure only)."""

"""Empty stand-in so `import flowcon` works without matplotlib (test infrastructure only)."""

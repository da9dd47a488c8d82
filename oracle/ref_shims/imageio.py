"""Import shim (test infrastructure): imageio is imported by the reference but never used on the path."""

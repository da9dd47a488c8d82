"""Import shim (test infrastructure) for solidspy; only uelutil.elast_quad4 is used by the reference."""

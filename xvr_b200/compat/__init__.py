"""Import shims (see compat/diffdrr)."""

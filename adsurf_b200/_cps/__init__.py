"""Shims with the reference's `_cps` module names."""

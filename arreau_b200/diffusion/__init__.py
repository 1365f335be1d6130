"""Mirrors of the reference's `diffusion/` modules on the hot path (same names and signatures)."""

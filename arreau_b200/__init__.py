"""arreau_b200: B200-native (sm_100a) implementation of the Arreau denoising step behind the reference's
Python module API.  See DESIGN.md; every computation is a kernel of libarreau_b200.so (no fallback)."""
from .tables import DiffusionTables, build_tables  # noqa: F401

__all__ = ["DiffusionTables", "build_tables"]

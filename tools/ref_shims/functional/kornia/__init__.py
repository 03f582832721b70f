"""Stand-in for the slice of kornia that Free-SurGS executes (see tools/ref_shims/__init__.py)."""
from . import geometry  # noqa: F401

"""skimage stand-in: only ``skimage.metrics.structural_similarity`` (utils/general_utils.py:42)."""
from . import metrics  # noqa: F401

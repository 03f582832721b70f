from . import conversions, epipolar, linalg  # noqa: F401

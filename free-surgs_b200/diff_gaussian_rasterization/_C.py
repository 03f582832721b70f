"""Stand-in for the reference package's pybind module ``diff_gaussian_rasterization._C``:
same three function names; each forwards to the C-ABI library (``include/fsgs_raster.h``)."""
from fsgs_b200.rasterizer import mark_visible, rasterize_gaussians, rasterize_gaussians_backward  # noqa: F401
